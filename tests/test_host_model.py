"""CPU tests of the host-side logic: the product's host ViT + apla helpers reproduce the reference's construction
bit-exactly (weights, indices, key set, trainable set) as recorded in tests/golden/, the C-ABI library loads and
exports every symbol of include/apla_b200.h, and the product path refuses to run without CUDA."""
import hashlib
import os

import numpy as np
import pytest
import torch

from apla_b200.config import AplaConfig
from helpers import CASES, build_case


def digest(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


@pytest.mark.parametrize("name", list(CASES))
def test_construction_matches_reference(name):
    model, meta, arr = build_case(name)
    sd = model.state_dict()
    assert sorted(sd.keys()) == sorted(meta["state_keys"])                       # I6
    for k, d in meta["weights_digest"].items():
        assert digest(sd[k]) == d, f"{k} differs from the reference"             # I3 (bit-exact RNG replay)
    for k in sd:
        if k.endswith(".inds"):
            assert np.array_equal(sd[k].numpy().astype(np.int16), arr["inds/" + k])
    assert [n for n, p in model.named_parameters() if p.requires_grad] == meta["trainable"]   # I5 / I7


def test_build_apla_error_conventions():
    from apla_b200.apla import build_apla
    from apla_b200.hostvit import HostViT, VitArch
    vit = HostViT(VitArch(128, 1, 2), img_size=28, patch_size=14)
    with pytest.raises(AssertionError):
        build_apla(AplaConfig(16), vit, "apla_attn", is_multi_gpu=True)          # apla_vit.py:77
    with pytest.raises(NotImplementedError):
        build_apla(AplaConfig(16), vit, "something_else")                        # apla_vit.py:84-89
    cfg = AplaConfig(16, inds_path="x.json")
    assert hasattr(cfg, "inds_path") and "inds_path" in cfg and not hasattr(AplaConfig(16), "inds_path")


def test_module_surface():
    """Attributes / parameters / buffers of APLA_Attention named as in appla_attn.py:11-48; MemEff subclass identity."""
    from apla_b200.apla.appla_attn import APLA_Attention
    from apla_b200.apla.appla_attn_mem_eff import APLA_MemEffAttention
    torch.manual_seed(0)
    m = APLA_MemEffAttention(AplaConfig(8), dim=768, num_heads=12, qkv_bias=True)
    assert isinstance(m, APLA_Attention)
    assert m.inds[:16].tolist() == [428, 757, 549, 648, 587, 161, 271, 672, 625, 303, 579, 399, 720, 492, 622, 704]
    assert set(dict(m.named_parameters())) == {"proj_weight1", "proj_weight2", "proj_bias1", "proj_bias2", "qkv.weight",
                                               "qkv.bias"}
    assert m.proj_weight1.shape == (8, 768) and m.proj_weight2.shape == (760, 768)
    assert m.proj_weight1.requires_grad and m.proj_bias1.requires_grad
    assert not (m.proj_weight2.requires_grad or m.qkv.weight.requires_grad)
    assert torch.equal(m.trainable_inds, m.inds[:8]) and torch.equal(m.freezed_inds, m.inds[8:])
    for attr in ("num_heads", "scale", "partial_size", "dim", "indices", "attn_drop", "proj_drop"):
        assert hasattr(m, attr)


def test_no_cpu_fallback():
    from apla_b200.apla.appla_attn import APLA_Attention
    m = APLA_Attention(AplaConfig(8), dim=128, num_heads=2, qkv_bias=True)
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 4, 128))


def test_library_exports_every_declared_symbol():
    from apla_b200._lib import LIB, LIB_PATH
    if not os.path.exists(LIB_PATH):
        from apla_b200.build import build
        build()
    dll = LIB.load()
    assert len(LIB.protos) >= 30
    for name in LIB.protos:
        assert hasattr(dll, name)
    assert dll.apla_version() >= 100
    # engine argument validation needs no GPU
    assert not dll.apla_engine_create(2, 17, 100, 2, 2, 512, 10, 14, 56, 640, 16, 64, 0, 1e-6, 0.125)
    assert "invalid configuration" in LIB.last_error()


def test_varlen_mask_helper():
    from apla_b200.apla.appla_attn_mem_eff import BlockDiagonalMask, _seqlens_of
    m = BlockDiagonalMask.from_seqlens([257, 257, 50])
    assert m.cu_seqlens("cpu").tolist() == [0, 257, 514, 564] and _seqlens_of(m) == [257, 257, 50]

    class FakeX:      # duck-typed xformers mask
        class q_seqinfo:
            seqstart_py = [0, 3, 10]
    assert _seqlens_of(FakeX()) == [3, 7]
    with pytest.raises(AssertionError):
        _seqlens_of(object())


def test_fuse_apla_blocks_host_logic():
    """Block-level drop-in, the part that needs no GPU: the swap keeps every state-dict key and parameter object, is
    idempotent, refuses models without APLA attention, and the fused block has no CPU path."""
    import ctypes
    from apla_b200._lib import LIB, BlockWeights
    from apla_b200.apla import FusedAplaBlock, fuse_apla_blocks
    from apla_b200.hostvit import VitArch, build_classifier
    model, meta, _ = build_case("tiny_r16")
    keys = list(model.state_dict().keys())
    params = {n: id(p) for n, p in model.named_parameters()}
    assert fuse_apla_blocks(model) is model
    assert all(isinstance(b, FusedAplaBlock) for b in model.backbone.blocks)
    assert list(model.state_dict().keys()) == keys == meta["state_keys"]
    assert {n: id(p) for n, p in model.named_parameters()} == params
    assert [n for n, p in model.named_parameters() if p.requires_grad] == meta["trainable"]
    fuse_apla_blocks(model)                                                       # second call: nothing left to swap
    assert list(model.state_dict().keys()) == keys
    model.float()                                                                 # nn.Module._apply still reaches the children
    with pytest.raises(RuntimeError, match="CUDA"):
        model.backbone.blocks[0](torch.randn(1, 17, 128))
    with pytest.raises(RuntimeError, match="CUDA"):
        model.backbone.blocks[0]([torch.randn(1, 17, 128), torch.randn(2, 5, 128)])
    with pytest.raises(AssertionError):
        model.backbone.blocks[0]([])
    stock = build_classifier(VitArch(128, 2, 2), img_size=56, patch_size=14, n_classes=10,
                             apla_config=AplaConfig("full"), is_multi_gpu=True, seed=0)
    with pytest.raises(RuntimeError, match="APLA_Attention"):
        fuse_apla_blocks(stock)
    with pytest.raises(AttributeError):
        fuse_apla_blocks(torch.nn.Linear(2, 2))
    # the ctypes mirror of apla_block_weights is the size the library was compiled with
    assert LIB.load().apla_block_weights_size() == ctypes.sizeof(BlockWeights)


def test_cache_pos_encoding_is_transparent():
    """cache_pos_encoding: same values as the un-cached resize, computed once per grid, recomputed when the table changes,
    and never cached while the table requires gradients."""
    import torch
    from apla_b200.apla import cache_pos_encoding
    from apla_b200.hostvit import HostViT, VitArch
    torch.manual_seed(0)
    vit = HostViT(VitArch(128, 2, 2), img_size=56, patch_size=14)
    vit.pos_embed.requires_grad = False
    want = vit.pos_for(4).clone()
    cache_pos_encoding(vit)
    a, b = vit.pos_for(4), vit.pos_for(4)
    assert a is b and torch.equal(a, want)
    assert vit.pos_for(16) is vit.pos_embed or torch.equal(vit.pos_for(16), vit.pos_embed)      # native grid: the table itself
    with torch.no_grad():
        vit.pos_embed.mul_(2.0)                                   # in-place update bumps the version: cache miss
    assert torch.allclose(vit.pos_for(4), 2 * want, atol=1e-6)
    vit.pos_embed.requires_grad = True
    assert vit.pos_for(4).requires_grad                           # trainable table: the live computation, with autograd
