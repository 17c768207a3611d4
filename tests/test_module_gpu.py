"""Module-level drop-in on the B200: `APLA_Attention` / `APLA_MemEffAttention` from apla_b200.apla, called the way the
reference's ViT blocks call them (src/utils/transformers/vit.py:280, dinov2/layers/block.py:253-288), against an fp32
torch restatement of the reference forward (src/apla/appla_attn.py:50-83: qkv Linear, softmax attention, two F.linear
on the trainable / frozen row sets, two scatter_).  Checks the output and the gradients autograd receives: input
gradient, and weight / bias gradient of the trainable rows ONLY (frozen tensors get none).
Bars: relative error <= 1e-2, gradient cosine >= 0.999 (BASELINE.json north_star)."""
import pytest
import torch

from helpers import cosine, rel

pytestmark = pytest.mark.gpu


def _mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from apla_b200.apla import APLA_Attention, APLA_MemEffAttention
    from apla_b200.apla.appla_attn_mem_eff import BlockDiagonalMask
    from apla_b200.config import AplaConfig
    return APLA_Attention, APLA_MemEffAttention, BlockDiagonalMask, AplaConfig


def _init(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() == 2 else 0.1))
    return mod.cuda()


def _reference(mod, x, seqlens):
    """fp32 restatement of appla_attn.py:50-83 (dense) / appla_attn_mem_eff.py:37-67 (block-diagonal) on the module's
    own parameters; x [B,N,C] or packed [1, sum(N), C]."""
    H, C = mod.num_heads, mod.dim
    w1 = mod.proj_weight1.detach().float().clone().requires_grad_(True)
    b1 = mod.proj_bias1.detach().float().clone().requires_grad_(True)
    xr = x.detach().float().clone().requires_grad_(True)
    qkv = torch.nn.functional.linear(xr, mod.qkv.weight.float(), mod.qkv.bias.float() if mod.qkv.bias is not None else None)
    flat = qkv.reshape(-1, 3, H, C // H)
    outs, o = [], 0
    for n in seqlens:
        t = flat[o:o + n].permute(1, 2, 0, 3)                    # [3, H, n, hd]
        a = ((t[0] @ t[1].transpose(-2, -1)) * mod.scale).softmax(-1)
        outs.append((a @ t[2]).transpose(0, 1).reshape(n, C))
        o += n
    ao = torch.cat(outs, 0)
    out = torch.empty(ao.shape[0], C, device=x.device)
    tr, fr = mod.trainable_inds.to(x.device), mod.freezed_inds.to(x.device)
    out[:, tr] = torch.nn.functional.linear(ao, w1, b1)
    if fr.numel():
        out[:, fr] = torch.nn.functional.linear(ao, mod.proj_weight2.float(), mod.proj_bias2.float())
    return out.view(x.shape), xr, w1, b1


@pytest.mark.parametrize("B,N,dim,heads,r", [(3, 257, 128, 2, 16), (2, 197, 384, 6, 32), (2, 50, 256, 4, 256)])
def test_apla_attention_forward_backward(B, N, dim, heads, r):
    APLA_Attention, _, _, AplaConfig = _mods()
    torch.manual_seed(7)
    mod = _init(APLA_Attention(AplaConfig(r), dim, num_heads=heads, qkv_bias=True), seed=B * N + r)
    assert [n for n, p in mod.named_parameters() if p.requires_grad] == ["proj_weight1", "proj_bias1"]
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(B, N, dim, device="cuda", generator=g).requires_grad_(True)
    out, attn = mod(x)
    assert attn is None and out.shape == x.shape
    ref, xr, w1, b1 = _reference(mod, x, [N] * B)
    assert rel(out.detach().float(), ref.detach()) <= 1e-2
    dy = torch.randn(B, N, dim, device="cuda", generator=g)
    out.backward(dy)
    ref.backward(dy)
    for name, ours, theirs in (("dx", x.grad, xr.grad), ("dW1", mod.proj_weight1.grad, w1.grad),
                               ("db1", mod.proj_bias1.grad, b1.grad)):
        assert rel(ours.float(), theirs) <= 1.2e-2, (name, rel(ours.float(), theirs))
        assert cosine(ours.float().flatten().cpu(), theirs.flatten().cpu()) >= 0.999, name
    assert mod.proj_weight2.grad is None and mod.qkv.weight.grad is None          # frozen: no gradient allocated


def test_apla_mem_eff_attention_block_diagonal():
    """Packed crops of different length (dinov2 multi-crop: 257-token global + 50-token local crops) through the
    BlockDiagonalMask path == independent attention per crop."""
    _, APLA_MemEffAttention, BlockDiagonalMask, AplaConfig = _mods()
    torch.manual_seed(3)
    dim, heads, r = 256, 4, 256
    mod = _init(APLA_MemEffAttention(AplaConfig(r), dim, num_heads=heads, qkv_bias=True), seed=5)
    seqlens = [257, 257, 50, 50, 50, 50]
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(1, sum(seqlens), dim, device="cuda", generator=g).requires_grad_(True)
    out = mod(x, attn_bias=BlockDiagonalMask.from_seqlens(seqlens))
    assert isinstance(out, torch.Tensor) and out.shape == x.shape
    ref, xr, w1, b1 = _reference(mod, x, seqlens)
    assert rel(out.detach().float(), ref.detach()) <= 1e-2
    dy = torch.randn_like(out)
    out.backward(dy)
    ref.backward(dy)
    assert rel(x.grad.float(), xr.grad) <= 1.2e-2
    assert rel(mod.proj_weight1.grad, w1.grad) <= 1.2e-2
    assert cosine(mod.proj_weight1.grad.flatten().cpu(), w1.grad.flatten().cpu()) >= 0.999
    with pytest.raises(AssertionError):
        mod(x[:, :-1], attn_bias=BlockDiagonalMask.from_seqlens(seqlens))       # packed length must match the mask


def test_dropout_raises_and_cpu_input_raises():
    APLA_Attention, _, _, AplaConfig = _mods()
    mod = _init(APLA_Attention(AplaConfig(8), 128, num_heads=2, qkv_bias=True, proj_drop=0.1), seed=1)
    mod.train()
    with pytest.raises(RuntimeError):
        mod(torch.randn(1, 17, 128, device="cuda"))
    mod2 = _init(APLA_Attention(AplaConfig(8), 128, num_heads=2, qkv_bias=True), seed=1)
    with pytest.raises(RuntimeError):
        mod2(torch.randn(1, 17, 128))                                              # no CPU fallback


def test_step_is_deterministic():
    """Same inputs, same state: logits and the saved residual stream are bit-identical between runs; gradients move
    only by the fp32 atomics of the split-K weight gradient."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from apla_b200.config import AplaConfig
    from apla_b200.engine import FineTuneEngine
    from apla_b200.hostvit import VitArch, build_classifier
    model = build_classifier(VitArch(128, 3, 2), img_size=224, patch_size=14, n_classes=37, apla_config=AplaConfig(16), seed=0)
    eng = FineTuneEngine(model, batch_size=12, img_size=224, device="cuda:0")
    g = torch.Generator().manual_seed(9)
    images = torch.randn(12, 3, 224, 224, generator=g).cuda()
    labels = torch.randint(0, 37, (12,), generator=g).cuda()
    runs = []
    for _ in range(3):
        eng.forward(images, labels)
        eng.backward()
        torch.cuda.synchronize()
        runs.append((eng.logits.clone(), eng.xs[-1].clone(), eng.grads.clone()))
    for lg, xs, gr in runs[1:]:
        assert torch.equal(lg, runs[0][0]) and torch.equal(xs, runs[0][1])
        assert float((gr - runs[0][2]).norm() / runs[0][2].norm()) < 1e-5


def test_refresh_working_set_after_a_write_through_data():
    """The bf16 working copies are keyed on (data_ptr, _version): an optimiser step or copy_ is noticed, a write through
    `.data` is not -- `refresh_working_set()` is the documented way to pick it up."""
    APLA_Attention, _, _, AplaConfig = _mods()
    mod = _init(APLA_Attention(AplaConfig(8), 128, num_heads=2, qkv_bias=True), seed=4).eval()
    x = torch.randn(2, 33, 128, device="cuda")
    with torch.no_grad():
        y0, _ = mod(x)
        mod.proj_weight1.add_(0.5)                               # versioned write: seen
        y1, _ = mod(x)
        assert rel(y1, y0) > 1e-3
        mod.proj_weight1.data.add_(0.5)                          # bypasses the version counter: not seen ...
        y2, _ = mod(x)
        assert torch.equal(y2, y1)
        mod.refresh_working_set()                                # ... until asked
        y3, _ = mod(x)
        assert rel(y3, y1) > 1e-3
        mod.qkv.weight.data.mul_(0.5)                            # frozen tensors likewise
        mod.refresh_working_set()
        y4, _ = mod(x)
        assert rel(y4, y3) > 1e-3
