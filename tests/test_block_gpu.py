"""Block-level drop-in on the B200 (`apla_b200.apla.fuse_apla_blocks` / `FusedAplaBlock`): the whole `Block.forward`
(src/utils/transformers/vit.py:279-288) and the packed multi-crop `NestedTensorBlock.forward_nested`
(src/self_supervised/dinov2/layers/block.py:274-288) as one autograd node, against
 (a) the fp32 CPU oracle's `block_forward` on the same tensors (output, gradient w.r.t. EVERY input token, weight / bias
     gradient of the trainable projection rows), dense and block-diagonal, and
 (b) the golden vectors recorded from the unmodified reference for a whole model built from fused blocks.
Bars (BASELINE.json north_star): relative error <= 1e-2, gradient cosine >= 0.999."""
import pytest
import torch

from helpers import build_case, cosine, rel, synthetic_batch

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _oracle_block(model, l, x_cpu, seqlens=None):
    """oracle.block_forward on block l of `model` (CPU fp32, autograd on x / proj_weight1 / proj_bias1)."""
    from oracle import apla_oracle as O
    sd = {k: (v.detach().float() if v.is_floating_point() else v.detach()).cpu().clone()
          for k, v in model.state_dict().items()}
    b = f"backbone.blocks.{l}."
    w1 = sd[b + "attn.proj_weight1"].requires_grad_(True)
    b1 = sd[b + "attn.proj_bias1"].requires_grad_(True)
    xr = x_cpu.clone().requires_grad_(True)
    blk = model.backbone.blocks[l]
    out = O.block_forward(sd, b, xr, blk.attn.num_heads, float(blk.norm1.eps), seqlens)
    return out, xr, w1, b1


@pytest.mark.parametrize("name,l,B,N", [("tiny_r16", 1, 3, 257), ("c1_vits16_r32_pert", 5, 2, 197),
                                        ("tiny_interp_r128", 0, 2, 50), ("c5_vitb14_518_r768", 3, 1, 1370)])
def test_fused_block_matches_oracle_dense(name, l, B, N):
    _need_gpu()
    from apla_b200.apla import FusedAplaBlock, fuse_apla_blocks
    model, _, _ = build_case(name)
    D = model.backbone.embed_dim
    g = torch.Generator().manual_seed(11)
    x_cpu = torch.randn(B, N, D, generator=g)
    dy_cpu = torch.randn(B, N, D, generator=g)
    ref, xr, w1, b1 = _oracle_block(model, l, x_cpu)
    ref.backward(dy_cpu)

    keys = list(model.state_dict().keys())
    fuse_apla_blocks(model.cuda())
    assert list(model.state_dict().keys()) == keys                      # same checkpoint keys after the swap
    blk = model.backbone.blocks[l]
    assert isinstance(blk, FusedAplaBlock)
    x = x_cpu.cuda().requires_grad_(True)
    out = blk(x)
    assert out.shape == x.shape and out.dtype == x.dtype
    out.backward(dy_cpu.cuda())
    torch.cuda.synchronize()
    at = blk.attn
    assert rel(out.detach(), ref.detach()) <= 1e-2, rel(out.detach(), ref.detach())
    assert rel(x.grad, xr.grad) <= 1e-2 and cosine(x.grad, xr.grad) >= 0.999, rel(x.grad, xr.grad)
    assert rel(at.proj_weight1.grad, w1.grad) <= 1e-2, rel(at.proj_weight1.grad, w1.grad)
    assert cosine(at.proj_weight1.grad, w1.grad) >= 0.999
    assert rel(at.proj_bias1.grad, b1.grad) <= 1e-2
    # nothing else received a gradient
    for n_, p in blk.named_parameters():
        if "proj_weight1" not in n_ and "proj_bias1" not in n_:
            assert p.grad is None, n_
    # input that does not require grad (block 0 behind a frozen embedding): only the weight gradient is computed
    for p in blk.parameters():
        p.grad = None
    blk(x_cpu.cuda()).backward(dy_cpu.cuda())
    assert rel(at.proj_weight1.grad, w1.grad) <= 1e-2


@pytest.mark.parametrize("name,l,n_global,n_local", [("tiny_r16", 1, 2, 4),
                                                     ("vitl14_r128", 10, 2, 8)])      # C4: ViT-L, 2 global + 8 local crops
def test_fused_block_packed_crops(name, l, n_global, n_local):
    """dinov2 multi-crop: a list of [b_i, N_i, D] crop tensors is packed and attended block-diagonally; every crop's
    tokens get gradients (iBOT / dense losses), the projection rows get one summed weight gradient."""
    _need_gpu()
    from apla_b200.apla import fuse_apla_blocks
    model, _, _ = build_case(name)
    D = model.backbone.embed_dim
    g = torch.Generator().manual_seed(5)
    crops = [torch.randn(n_global, 257, D, generator=g), torch.randn(n_local, 50, D, generator=g)]
    dys = [torch.randn(c.shape, generator=g) for c in crops]
    seqlens = [257] * n_global + [50] * n_local
    packed = torch.cat([c.reshape(1, -1, D) for c in crops], 1)
    ref, xr, w1, b1 = _oracle_block(model, l, packed, seqlens)
    ref.backward(torch.cat([d.reshape(1, -1, D) for d in dys], 1))

    fuse_apla_blocks(model.cuda())
    blk = model.backbone.blocks[l]
    xs = [c.cuda().requires_grad_(True) for c in crops]
    outs = blk(xs)
    assert isinstance(outs, list) and [o.shape for o in outs] == [c.shape for c in crops]
    torch.autograd.backward(outs, [d.cuda() for d in dys])
    torch.cuda.synchronize()
    got = torch.cat([o.detach().reshape(1, -1, D) for o in outs], 1)
    assert rel(got, ref.detach()) <= 1e-2
    dx = torch.cat([x.grad.reshape(1, -1, D) for x in xs], 1)
    assert rel(dx, xr.grad) <= 1e-2 and cosine(dx, xr.grad) >= 0.999
    assert rel(blk.attn.proj_weight1.grad, w1.grad) <= 1e-2 and cosine(blk.attn.proj_weight1.grad, w1.grad) >= 0.999
    assert rel(blk.attn.proj_bias1.grad, b1.grad) <= 1e-2


def test_packed_crop_list_is_passed_on_without_repacking():
    """Two fused blocks in a row: the PackedCropList the first one returns goes into the second as ONE packed tensor (no
    torch.cat), and gives the same outputs and the same gradients as a plain list of the same crops (bitwise: same
    kernels, same inputs)."""
    _need_gpu()
    from apla_b200.apla import fuse_apla_blocks
    from apla_b200.apla.apla_block import PackedCropList
    model, _, _ = build_case("tiny_r16")
    D = model.backbone.embed_dim
    fuse_apla_blocks(model.cuda())
    b0, b1 = model.backbone.blocks[0], model.backbone.blocks[1]
    g = torch.Generator().manual_seed(6)
    crops = [torch.randn(2, 257, D, generator=g), torch.randn(3, 50, D, generator=g)]
    dys = [torch.randn(c.shape, generator=g).cuda() for c in crops]

    def run(repack):
        for p in model.parameters():
            p.grad = None
        xs = [c.cuda().requires_grad_(True) for c in crops]
        mid = b0(xs)
        assert isinstance(mid, PackedCropList) and mid.packed_if_untouched() is not None
        if repack:
            mid = [m for m in mid]                   # a plain list: the second block concatenates the views again
        outs = b1(mid)
        torch.autograd.backward(list(outs), dys)
        torch.cuda.synchronize()
        return [o.detach().clone() for o in outs], [x.grad.clone() for x in xs], b0.attn.proj_weight1.grad.clone()

    o1, g1, w1 = run(False)
    o2, g2, w2 = run(True)
    for a, b in zip(o1 + g1, o2 + g2):
        assert torch.equal(a, b)
    assert rel(w1, w2) < 1e-6                        # (split-K weight gradient: fp32 atomics)
    mid = b0([c.cuda() for c in crops])
    mid[0] = mid[0] * 1.0                            # a modified list must not use the stale packed tensor
    assert mid.packed_if_untouched() is None


@pytest.mark.parametrize("S,patch,D", [(224, 14, 768), (98, 14, 1024), (224, 16, 384)])
def test_fused_patch_embed_matches_conv(S, patch, D):
    """FusedPatchEmbed (apla_patchify + tcgen05 GEMM, bf16 operands) against the fp32 convolution it replaces."""
    _need_gpu()
    from apla_b200.apla import fuse_patch_embed
    from apla_b200.hostvit import PatchProjection

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.patch_embed = PatchProjection(S, patch, D)

    torch.manual_seed(3)
    m = Holder()
    for p in m.parameters():
        p.requires_grad = False
    x = torch.randn(5, 3, S, S)
    ref = m.patch_embed(x)
    keys = list(m.state_dict().keys())
    fuse_patch_embed(m.cuda())
    assert list(m.state_dict().keys()) == keys
    out = m.patch_embed(x.cuda())
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel(out, ref) <= 6e-3
    m.patch_embed.proj.weight.requires_grad = True
    with pytest.raises(RuntimeError):
        m.patch_embed(x.cuda())


@pytest.mark.parametrize("name", ["tiny_r16", "c1_vits16_r32_pert", "c2_vitb14_r8"])
def test_model_of_fused_blocks_matches_reference_golden(name):
    """The host model with fused blocks, driven by plain PyTorch autograd (no step engine): logits, loss and the
    gradients of every trainable tensor against the vectors recorded from the unmodified reference."""
    _need_gpu()
    from apla_b200.apla import fuse_apla_blocks
    model, meta, arr = build_case(name)
    m = meta["meta"]
    fuse_apla_blocks(model.cuda())
    images, labels = synthetic_batch(m["batch"], m["img"], m["n_classes"])
    logits = model(images.cuda())
    loss = torch.nn.functional.cross_entropy(logits, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert rel(logits.detach(), arr["s0/logits"]) <= 1e-2, rel(logits.detach(), arr["s0/logits"])
    assert abs(float(loss) - float(arr["s0/loss"])) <= 1e-2 * abs(float(arr["s0/loss"]))
    named = dict(model.named_parameters())
    assert [k for k, p in named.items() if p.requires_grad] == meta["trainable"]
    sub = m["sub"]
    ours = torch.cat([named[k].grad.flatten()[::sub].cpu() for k in meta["trainable"]])
    ref = torch.cat([torch.as_tensor(arr["s0/grad/" + k]).flatten() for k in meta["trainable"]])
    assert cosine(ours, ref) >= 0.999, cosine(ours, ref)
    assert rel(ours, ref) <= 1e-2, rel(ours, ref)


def test_fused_block_refuses_what_it_does_not_implement():
    _need_gpu()
    from apla_b200.apla import FusedAplaBlock, fuse_apla_blocks
    from apla_b200.config import AplaConfig
    from apla_b200.hostvit import VitArch, build_classifier
    model, _, _ = build_case("tiny_r16")
    fuse_apla_blocks(model)
    blk = model.backbone.blocks[0]
    with pytest.raises(RuntimeError, match="CUDA"):
        blk(torch.randn(1, 17, 128))                                    # no CPU fallback
    model.cuda()
    with pytest.raises(RuntimeError, match="probabilities"):
        blk(torch.randn(1, 17, 128, device="cuda"), return_attention=True)
    blk.norm1.weight.requires_grad_(True)
    with pytest.raises(RuntimeError, match="frozen"):
        blk(torch.randn(1, 17, 128, device="cuda"))
    blk.norm1.weight.requires_grad_(False)
    blk.attn.proj_drop.p = 0.1
    blk.train()
    with pytest.raises(RuntimeError, match="dropout"):
        blk(torch.randn(1, 17, 128, device="cuda"))
    # multi-GPU 'full' keeps the stock attention (apla_vit.py:65-75): nothing to fuse
    stock = build_classifier(VitArch(128, 2, 2), img_size=56, patch_size=14, n_classes=10,
                             apla_config=AplaConfig("full"), is_multi_gpu=True, seed=0)
    with pytest.raises(RuntimeError, match="APLA_Attention"):
        fuse_apla_blocks(stock)
    with pytest.raises(TypeError):
        FusedAplaBlock(stock.backbone.blocks[0])


def _chain_case(chain, extra_consumer, packed):
    """Three fused blocks in a row; gradients of the input and of every trainable row, with / without the block-to-block
    hand-over of bf16(gamma2 * dx), optionally with a SECOND consumer of an intermediate output (its gradient is then a sum
    autograd forms, which the hand-over must not be used for)."""
    import apla_b200.apla.apla_block as AB
    from apla_b200._lib import LIB
    from apla_b200.apla import fuse_apla_blocks
    from apla_b200.config import AplaConfig
    from apla_b200.hostvit import VitArch, build_classifier
    from helpers import perturb_module
    AB._CHAIN = chain
    model = build_classifier(VitArch(128, 3, 2), img_size=56, patch_size=14, n_classes=10, apla_config=AplaConfig(16), seed=0)
    perturb_module(model)                                        # LayerScale vectors away from 1, biases away from 0
    fuse_apla_blocks(model.cuda())
    blocks = list(model.backbone.blocks)
    D = model.backbone.embed_dim
    g = torch.Generator().manual_seed(5)
    if packed:
        xs = [torch.randn(2, 257, D, generator=g).cuda().requires_grad_(True), torch.randn(4, 50, D, generator=g).cuda().requires_grad_(True)]
        h, mids = xs, []
        for b in blocks:
            h = b(h)
            mids.append(h)
        loss = sum((t.float() ** 2).sum() for t in h)
        if extra_consumer:
            loss = loss + (mids[0][0].float() * 0.37).sum()
        leaves = xs
    else:
        x = torch.randn(3, 65, D, generator=g).cuda().requires_grad_(True)
        h, mids = x, []
        for b in blocks:
            h = b(h)
            mids.append(h)
        loss = (h.float() ** 2).sum()
        if extra_consumer:
            loss = loss + (mids[1].float() * 0.37).sum()
        leaves = [x]
    n0 = LIB.load().apla_launch_count()
    loss.backward()
    torch.cuda.synchronize()
    launches = int(LIB.load().apla_launch_count() - n0)
    grads = [t.grad.clone() for t in leaves]
    for b in blocks:
        grads += [b.attn.proj_weight1.grad.clone(), b.attn.proj_bias1.grad.clone()]
    AB._CHAIN = True
    return grads, launches, len(blocks)


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("extra_consumer", [False, True])
def test_gradient_hand_over_between_fused_blocks(packed, extra_consumer):
    _need_gpu()
    ref, l_ref, n = _chain_case(False, extra_consumer, packed)
    got, l_got, _ = _chain_case(True, extra_consumer, packed)
    for a, b in zip(got, ref):
        assert rel(a, b) < 1e-5, rel(a, b)                      # (weight gradients: fp32 atomics of the split-K kernel)
    assert torch.equal(got[0], ref[0])                           # input gradient: same kernels, same operands
    # every hand-over replaces one launch (the cast) of the block in front; the second consumer blocks it for one block
    saved = l_ref - l_got
    assert saved == (n - 1) - (1 if extra_consumer else 0), (l_ref, l_got)
