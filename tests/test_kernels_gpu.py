"""Kernel-level parity on the B200: every C-ABI kernel against a plain PyTorch fp32 restatement of the same op
on the same (bf16-rounded) inputs.  Tolerances are the bf16 output rounding (2^-8 relative) unless stated."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from apla_b200 import ops
    return ops


def rel(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 384, 128), (1576, 1152, 384), (16448, 768, 768),
                                   (514, 2304, 768), (2 * 257, 3072, 768), (100, 128, 3072), (257, 32, 64)])
def test_gemm_bias(M, N, K):
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    w = bf(torch.randn(N, K, device="cuda", generator=g) * 0.05)
    bias = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm_bias(a, w, bias)
    ref = a.float() @ w.float().t() + bias
    assert rel(out.float(), ref) < 4e-3
    assert torch.allclose(out.float(), ref, atol=2e-2, rtol=1e-2)
    out2 = ops.gemm_bias(a, w, None)
    assert rel(out2.float(), a.float() @ w.float().t()) < 4e-3


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_all_tile_widths(bn):
    """Force each BN instantiation through the internal override (exported for tests via env)."""
    ops = _cuda()
    import os
    os.environ["APLA_GEMM_BN"] = str(bn)
    try:
        g = torch.Generator(device="cuda").manual_seed(bn)
        a = bf(torch.randn(1000, 512, device="cuda", generator=g))
        w = bf(torch.randn(768, 512, device="cuda", generator=g) * 0.05)
        out = ops.gemm_bias(a, w, None)
        assert rel(out.float(), a.float() @ w.float().t()) < 4e-3
    finally:
        del os.environ["APLA_GEMM_BN"]


def test_gemm_bias_gelu():
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    a = bf(torch.randn(777, 384, device="cuda", generator=g))
    w = bf(torch.randn(1536, 384, device="cuda", generator=g) * 0.05)
    bias = torch.randn(1536, device="cuda", generator=g) * 0.1
    h, gl = ops.gemm_bias_gelu(a, w, bias)
    href = a.float() @ w.float().t() + bias
    assert rel(h.float(), href) < 4e-3
    gref = torch.nn.functional.gelu(h.float())          # GELU of the stored bf16 pre-activation
    assert rel(gl.float(), gref) < 4e-3


def test_gemm_ls_residual_inplace():
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 1028, 768, 3072
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    w = bf(torch.randn(N, K, device="cuda", generator=g) * 0.02)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = resid + gamma * (a.float() @ w.float().t() + bias)
    out = ops.gemm_bias_ls_residual(a, w, bias, gamma, resid)
    assert rel(out, ref) < 1e-3
    r2 = resid.clone()
    ops.gemm_bias_ls_residual(a, w, bias, None, r2, out=r2)
    assert rel(r2, resid + (a.float() @ w.float().t() + bias)) < 1e-3


@pytest.mark.parametrize("M,N,K", [(16448, 768, 768), (1028, 768, 3072), (4099, 1024, 1024), (300, 384, 384), (64, 768, 768),
                                   (2000, 512, 256)])
def test_gemm_ls_residual_layernorm_one_launch(M, N, K):
    """The residual GEMM that also normalises its output (the slab's last column tile triggers the LayerNorm) against the
    two launches it replaces: identical residual stream, LayerNorm output equal to apla_layernorm_fwd of it -- repeated,
    because the hand-over between CTAs is a race if it is wrong."""
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    w = bf(torch.randn(N, K, device="cuda", generator=g) * 0.03)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) * 2 + 0.3
    ln_w = torch.randn(N, device="cuda", generator=g)
    ln_b = torch.randn(N, device="cuda", generator=g) * 0.1
    x_ref = ops.gemm_bias_ls_residual(a, w, bias, gamma, resid)
    y_ref = ops.layernorm_fwd(x_ref, ln_w, ln_b, 1e-6)
    want = torch.nn.functional.layer_norm(x_ref, (N,), ln_w, ln_b, 1e-6)
    assert rel(y_ref.float(), want) < 4e-3
    for it in range(6):
        x = torch.full((M, N), float("nan"), device="cuda")
        y = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm_bias_ls_residual_ln(a, w, bias, gamma, resid, ln_w, ln_b, 1e-6, out=x, ln_out=y, one_launch=1)
        assert torch.equal(x, x_ref), f"residual stream differs (iteration {it})"
        assert torch.equal(y, y_ref), f"LayerNorm output differs (iteration {it}): {rel(y.float(), y_ref.float()):.3e}"


def test_gemm_ls_accumulate_in_place():
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(21)
    M, N, K = 1028, 768, 768
    a = bf(torch.randn(M, K, device="cuda", generator=g))
    w = bf(torch.randn(N, K, device="cuda", generator=g) * 0.02)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    gamma = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = resid + gamma * (a.float() @ w.float().t() + bias)
    out = ops.gemm_bias_ls_accumulate(a, w, bias, gamma, resid.clone())
    assert rel(out, ref) < 1e-3


def test_gemm_dgrad_and_gelu_bwd():
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    M, Din, Dout = 900, 1536, 384           # fc2: in=1536, out=384
    dy = bf(torch.randn(M, Dout, device="cuda", generator=g))
    W = bf(torch.randn(Dout, Din, device="cuda", generator=g) * 0.05)    # [out, in]
    wt = W.t().contiguous()                                               # [in, out]
    dx = ops.gemm_dgrad(dy, wt)
    ref = dy.float() @ W.float()
    assert rel(dx.float(), ref) < 4e-3
    h = bf(torch.randn(M, Din, device="cuda", generator=g))
    dh = ops.gemm_dgrad_gelu_bwd(dy, wt, h)
    hh = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hh).backward(ref)
    assert rel(dh.float(), hh.grad) < 5e-3


def test_gemm_gelu_saved_derivative():
    """fc1 forward saving gelu'(h) in fp16 + fc2 dgrad multiplying by it == autograd through nn.GELU (fp32 reference)."""
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(11)
    M, Din, Dh = 900, 384, 1536
    a = bf(torch.randn(M, Din, device="cuda", generator=g))
    w1 = bf(torch.randn(Dh, Din, device="cuda", generator=g) * 0.08)
    b1 = torch.randn(Dh, device="cuda", generator=g) * 0.3
    d, gl = ops.gemm_bias_gelu_dgelu(a, w1, b1)
    h = (a.float() @ w1.float().t() + b1).requires_grad_(True)
    gref = torch.nn.functional.gelu(h)
    assert rel(gl.float(), gref.detach()) < 4e-3
    dy = bf(torch.randn(M, Din, device="cuda", generator=g))
    W2 = bf(torch.randn(Din, Dh, device="cuda", generator=g) * 0.05)      # fc2 weight [out, in]
    dh = ops.gemm_dgrad_mul(dy, W2.t().contiguous(), d)
    gref.backward(dy.float() @ W2.float())
    assert rel(d.float(), torch.autograd.grad(torch.nn.functional.gelu(h).sum(), h)[0]) < 1e-3
    assert rel(dh.float(), h.grad) < 4e-3


@pytest.mark.parametrize("M,D", [(514, 768), (1576, 384), (16448, 768)])
def test_gemm_dgrad_delta(M, D):
    """Projection dgrad with the fused delta = rowsum(dO * O) per (token, head) epilogue."""
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(M + D)
    dy = bf(torch.randn(M, D, device="cuda", generator=g))
    W = bf(torch.randn(D, D, device="cuda", generator=g) * 0.05)          # [out, in]
    o = bf(torch.randn(M, D, device="cuda", generator=g))
    d_o, delta = ops.gemm_dgrad_delta(dy, W.t().contiguous(), o)
    ref = dy.float() @ W.float()
    assert rel(d_o.float(), ref) < 4e-3
    assert torch.equal(d_o, ops.gemm_dgrad(dy, W.t().contiguous()))       # same GEMM, bit-identical dO
    dref = (d_o.float() * o.float()).view(M, D // 64, 64).sum(-1)          # delta of the ROUNDED dO
    assert rel(delta, dref) < 1e-5


@pytest.mark.parametrize("T,D,r", [(514, 768, 8), (1576, 384, 32), (2000, 768, 128), (16448, 768, 8)])
def test_proj_wgrad_compact(T, D, r):
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(T + r)
    x = bf(torch.randn(T, D, device="cuda", generator=g))
    dy = bf(torch.randn(T, D, device="cuda", generator=g))
    idx = torch.randperm(D, device="cuda", generator=g)[:r].to(torch.int32)
    n_pad = (r + 63) // 64 * 64
    sub = ops.gather_cols(dy, idx, r, n_pad)
    assert torch.equal(sub[:, :r], dy[:, idx.long()]) and float(sub[:, r:].abs().sum()) == 0.0
    dw = torch.zeros(r, D, device="cuda")
    ops.proj_wgrad(sub, x, dw, r)
    ref = dy[:, idx.long()].float().t() @ x.float()
    assert rel(dw, ref) < 1e-3
    db = torch.zeros(r, device="cuda")
    ops.colsum(sub, db, r)
    assert rel(db, dy[:, idx.long()].float().sum(0)) < 1e-3


def test_proj_wgrad_full_rowmap():
    """partial_size == dim: full dY as the B operand, rows permuted through rowmap in the epilogue."""
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(11)
    T, D = 1542, 768
    x = bf(torch.randn(T, D, device="cuda", generator=g))
    dy = bf(torch.randn(T, D, device="cuda", generator=g))
    perm = torch.randperm(D, device="cuda", generator=g)
    rowmap = torch.empty(D, dtype=torch.int32, device="cuda")
    rowmap[perm] = torch.arange(D, dtype=torch.int32, device="cuda")      # slot of output feature n
    dw = torch.zeros(D, D, device="cuda")
    ops.proj_wgrad(dy, x, dw, D, rowmap=rowmap)
    ref = dy[:, perm].float().t() @ x.float()
    assert rel(dw, ref) < 1e-3
    db = torch.zeros(D, device="cuda")
    ops.colsum(dy, db, D, rowmap=rowmap)
    assert rel(db, dy[:, perm].float().sum(0)) < 1e-3


@pytest.mark.parametrize("T,n,ld", [(16448, 768, 768), (58496, 1024, 1024), (1031, 768, 768), (7, 256, 256), (515, 264, 272),
                                    (300, 64, 64)])
def test_colsum_wide_and_narrow(T, n, ld):
    """Bias gradient = column sums of a bf16 matrix, scattered through rowmap (entries < 0 are dropped): the 16-byte-load
    kernel for wide matrices (n >= 256) and the narrow one, incl. a padded leading dimension and ragged row counts."""
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(T + n)
    full = bf(torch.randn(T, ld, device="cuda", generator=g))
    dy = full[:, :n]
    perm = torch.randperm(n, device="cuda", generator=g)
    rowmap = torch.empty(n, dtype=torch.int32, device="cuda")
    rowmap[perm] = torch.arange(n, dtype=torch.int32, device="cuda")
    rowmap[perm[: n // 5]] = -1                                           # a fifth of the columns has no slot
    db = torch.zeros(n, device="cuda")
    ops.colsum(dy, db, n, rowmap=rowmap)
    ref = torch.zeros(n, device="cuda")
    keep = rowmap >= 0
    ref[rowmap[keep].long()] = dy.float().sum(0)[keep]
    assert rel(db, ref) < 1e-4, rel(db, ref)
    db2 = torch.zeros(n, device="cuda")
    ops.colsum(dy, db2, n)
    assert rel(db2, dy.float().sum(0)) < 1e-4


@pytest.mark.parametrize("D", [128, 384, 768, 1024])
def test_layernorm_fwd_bwd(D):
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(D)
    rows, r = 1031, 16
    x = torch.randn(rows, D, device="cuda", generator=g) * 2 + 0.5
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.randn(D, device="cuda", generator=g)
    y = ops.layernorm_fwd(x, w, b, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6)
    assert rel(y.float(), ref) < 3e-3
    dy = bf(torch.randn(rows, D, device="cuda", generator=g))
    dres = torch.randn(rows, D, device="cuda", generator=g)
    gamma = torch.randn(D, device="cuda", generator=g)
    idx = torch.randperm(D, device="cuda", generator=g)[:r].to(torch.int32)
    xx = x.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xx, (D,), w, b, 1e-6).backward(dy.float())
    dx_ref = dres + xx.grad
    dxb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    sub = torch.empty(rows, 64, device="cuda", dtype=torch.bfloat16)
    dx = ops.layernorm_bwd(dy, x, w, 1e-6, dres=dres, dxb=dxb, gamma=gamma, sub=sub, idx=idx, r=r)
    assert rel(dx, dx_ref) < 1e-5
    assert rel(dxb.float(), gamma * dx_ref) < 3e-3
    assert rel(sub[:, :r].float(), (gamma * dx_ref)[:, idx.long()]) < 3e-3
    assert float(sub[:, r:].abs().sum()) == 0.0
    # in place on the residual gradient, no extras
    d2 = dres.clone()
    ops.layernorm_bwd(dy, x, w, 1e-6, dres=d2, dx=d2)
    assert rel(d2, dx_ref) < 1e-5
    # strided rows (CLS-only final norm): every 5th row of x
    ys = ops.layernorm_fwd(x[::5], w, b, 1e-6)
    assert rel(ys.float(), ref[::5]) < 3e-3


def _attn_ref(qkv, seqlens, H, scale):
    """fp32 restatement of appla_attn.py:53-60 per sequence, with autograd for the backward."""
    outs = []
    o = 0
    D = H * 64
    for n in seqlens:
        t = qkv[o:o + n].reshape(n, 3, H, 64).permute(1, 2, 0, 3)
        q, k, v = t[0], t[1], t[2]
        a = ((q @ k.transpose(-2, -1)) * scale).softmax(-1)
        outs.append((a @ v).transpose(0, 1).reshape(n, D))
        o += n
    return torch.cat(outs, 0)


@pytest.mark.parametrize("B,N,H", [(2, 257, 12), (30, 257, 12), (3, 197, 6), (1, 1370, 12), (4, 50, 16), (2, 64, 2), (5, 17, 2),
                                   (2, 272, 2)])
def test_attention_dense(B, N, H):
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(B * N + H)
    D = H * 64
    qkv = bf(torch.randn(B * N, 3 * D, device="cuda", generator=g))
    scale = 64 ** -0.5
    out, lse = ops.attn_fwd(qkv, H, scale, B, N)
    q32 = qkv.float().requires_grad_(True)
    ref = _attn_ref(q32, [N] * B, H, scale)
    assert rel(out.float(), ref) < 6e-3
    dout = bf(torch.randn(B * N, D, device="cuda", generator=g))
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, H, scale, B, N)
    assert rel(dqkv.float(), q32.grad) < 1e-2
    for part, name in ((slice(0, D), "dq"), (slice(D, 2 * D), "dk"), (slice(2 * D, 3 * D), "dv")):
        assert rel(dqkv[:, part].float(), q32.grad[:, part]) < 1.2e-2, name


@pytest.mark.parametrize("B,N,H", [(3, 257, 12), (2, 197, 6), (2, 1370, 2), (4, 50, 4)])
def test_attention_cls_only(B, N, H):
    """Last-block attention for the CLS query only == the dense kernels' CLS rows, and == their dqkv for a dout that is
    zero on every other row (fp32 torch reference)."""
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(B + N + H)
    D = H * 64
    qkv = bf(torch.randn(B * N, 3 * D, device="cuda", generator=g))
    scale = 0.125
    out, lse = ops.attn_cls_fwd(qkv, H, scale, B, N)
    q32 = qkv.float().requires_grad_(True)
    ref = _attn_ref(q32, [N] * B, H, scale)
    cls = torch.arange(B, device="cuda") * N
    assert rel(out[cls].float(), ref[cls]) < 6e-3
    assert float(out.float().abs().sum() - out[cls].float().abs().sum()) == 0.0      # other rows untouched (zero)
    dout = torch.zeros(B * N, D, device="cuda", dtype=torch.bfloat16)
    dout[cls] = bf(torch.randn(B, D, device="cuda", generator=g))
    ref.backward(dout.float())
    dqkv = ops.attn_cls_bwd(qkv, out, dout, lse, H, scale, B, N)
    for part, name in ((slice(0, D), "dq"), (slice(D, 2 * D), "dk"), (slice(2 * D, 3 * D), "dv")):
        assert rel(dqkv[:, part].float(), q32.grad[:, part]) < 1e-2, name
    mask = torch.ones(B * N, dtype=torch.bool, device="cuda")
    mask[cls] = False
    assert float(dqkv[mask][:, :D].float().abs().max()) == 0.0                          # dQ of non-CLS rows: exact zeros


@pytest.mark.parametrize("H,seqlens", [(4, [257, 50, 257, 50, 50, 3, 130]),
                                       # the DINOv2 multi-crop pattern: global crops, then MANY single-chunk local crops per
                                       # CTA (the second-generation backward once deadlocked on exactly this)
                                       (16, [257] * 6 + [50] * 90),
                                       (2, [257, 65, 256, 1, 129, 64, 257, 257, 200, 193] * 16)])
def test_attention_varlen(H, seqlens):
    ops = _cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    D = H * 64
    T = sum(seqlens)
    cu = torch.tensor([0] + list(torch.tensor(seqlens).cumsum(0)), dtype=torch.int32, device="cuda")
    qkv = bf(torch.randn(T, 3 * D, device="cuda", generator=g))
    scale = 0.125
    out, lse = ops.attn_fwd(qkv, H, scale, len(seqlens), max(seqlens), cu_seqlens=cu)
    q32 = qkv.float().requires_grad_(True)
    ref = _attn_ref(q32, seqlens, H, scale)
    assert rel(out.float(), ref) < 6e-3
    dout = bf(torch.randn(T, D, device="cuda", generator=g))
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, H, scale, len(seqlens), max(seqlens), cu_seqlens=cu)
    assert rel(dqkv.float(), q32.grad) < 1e-2


def test_clip_adamw_arena_matches_torch():
    """grad_sumsq + adamw_step over an arena == clip_grad_norm_ + torch.optim.AdamW (two groups), fp32."""
    _cuda()
    from apla_b200._lib import LIB, ptr, stream
    g = torch.Generator(device="cuda").manual_seed(8)
    n, n_decay = 100_003, 90_000
    p0 = torch.randn(n, device="cuda", generator=g) * 0.02
    ours = p0.clone(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    ss = torch.zeros(600, device="cuda")                      # APLA_SUMSQ_FLOATS: [0] result, rest scratch
    pa = torch.nn.Parameter(p0[:n_decay].clone()); pb = torch.nn.Parameter(p0[n_decay:].clone())
    opt = torch.optim.AdamW([{"params": [pa]}, {"params": [pb], "weight_decay": 0.0}], lr=3e-5, weight_decay=1e-5)
    for step in range(1, 4):
        grad = torch.randn(n, device="cuda", generator=g) * (3.0 if step == 1 else 1e-3)   # clipped / not clipped
        world = 2.0
        LIB.call("apla_grad_sumsq", ptr(grad), n, 1.0 / world, ptr(ss), stream())
        LIB.call("apla_adamw_step", ptr(ours), ptr(grad), ptr(m), ptr(v), n, n_decay, ptr(ss), 1.0 / world, 1.0, 3e-5,
                 1e-5, 0.9, 0.999, 1e-8, step, stream())
        pa.grad = grad[:n_decay] / world; pb.grad = grad[n_decay:] / world
        gn = torch.nn.utils.clip_grad_norm_([pa, pb], 1.0)
        opt.step()
        assert abs(float(ss[0].sqrt()) - float(gn)) <= 1e-5 * float(gn)
        assert float(ss[1:].abs().max()) >= 0.0 and int(ss[593:].view(torch.int32)[0]) == 0   # counter back at zero
        ref = torch.cat([pa.detach(), pb.detach()])
        assert rel(ours - p0, ref - p0) < 1e-4
        assert rel(ours, ref) < 1e-6


def test_head_and_cross_entropy():
    _cuda()
    from apla_b200._lib import LIB, ptr, stream
    g = torch.Generator(device="cuda").manual_seed(9)
    B, D, C = 16, 384, 555
    xn = bf(torch.randn(B, D, device="cuda", generator=g))
    W = torch.randn(C, D, device="cuda", generator=g) * 0.05
    bias = torch.randn(C, device="cuda", generator=g) * 0.1
    labels = torch.randint(0, C, (B,), device="cuda", generator=g)
    logits = torch.empty(B, C, device="cuda"); dlog = torch.empty(B, C, device="cuda"); loss = torch.zeros(1, device="cuda")
    LIB.call("apla_head_fwd", ptr(xn), ptr(W), ptr(bias), ptr(logits), B, D, C, stream())
    LIB.call("apla_cross_entropy", ptr(logits), ptr(labels), ptr(dlog), ptr(loss), B, C, 1.0 / B, 1.0 / B, stream())
    x32 = xn.float().requires_grad_(True); W32 = W.clone().requires_grad_(True); b32 = bias.clone().requires_grad_(True)
    ref_logits = x32 @ W32.t() + b32
    ref_loss = torch.nn.functional.cross_entropy(ref_logits, labels)
    ref_loss.backward()
    assert rel(logits, ref_logits) < 1e-5 and abs(float(loss) - float(ref_loss)) < 1e-5 * float(ref_loss)
    dW = torch.empty(C, D, device="cuda"); db = torch.empty(C, device="cuda")
    dxn = torch.empty(B, D, device="cuda", dtype=torch.bfloat16)
    LIB.call("apla_head_bwd", ptr(dlog), ptr(xn), ptr(W), ptr(dW), ptr(db), ptr(dxn), B, D, C, stream())
    assert rel(dW, W32.grad) < 1e-5 and rel(db, b32.grad) < 1e-5 and rel(dxn.float(), x32.grad) < 4e-3


def test_patchify_assemble_match_conv():
    _cuda()
    from apla_b200._lib import LIB, ptr, stream
    g = torch.Generator(device="cuda").manual_seed(10)
    B, S, p, D = 3, 56, 14, 128
    P, kk, kpad = (S // p) ** 2, 3 * p * p, 640
    img = torch.randn(B, 3, S, S, device="cuda", generator=g)
    patches = torch.empty(B * P, kpad, device="cuda", dtype=torch.bfloat16)
    LIB.call("apla_patchify", ptr(img), ptr(patches), B, S, p, kpad, stream())
    ref = torch.nn.functional.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(B * P, kk)
    assert torch.equal(patches[:, :kk], ref.to(torch.bfloat16)) and float(patches[:, kk:].abs().sum()) == 0.0
    pe = bf(torch.randn(B * P, D, device="cuda", generator=g))
    cls = torch.randn(D, device="cuda", generator=g); pos = torch.randn(P + 1, D, device="cuda", generator=g)
    x = torch.empty(B, P + 1, D, device="cuda")
    LIB.call("apla_assemble_tokens", ptr(pe), ptr(cls), ptr(pos), ptr(x), B, P, D, stream())
    want = torch.cat((cls.expand(B, 1, D), pe.float().view(B, P, D)), 1) + pos
    assert torch.equal(x, want)
