"""Per-kernel timing on the B200 (CUDA events, L2 flushed between iterations).  Prints one line per kernel with
achieved TFLOP/s or GB/s; used to fill DESIGN.md / profiles/, never as the bench.py number."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    T, D = 16448, 768
    res = []
    x = torch.randn(T, D, device=dev).bfloat16()
    x4 = torch.randn(T, 4 * D, device=dev).bfloat16()
    for name, N, K in (("qkv", 3 * D, D), ("proj", D, D), ("fc1", 4 * D, D), ("fc2", D, 4 * D)):
        a = x if K == D else x4
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        bias = torch.zeros(N, device=dev)
        out = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
        t = timeit(lambda: ops.gemm_bias(a, w, bias, out=out))
        res.append((f"gemm_bias {name} {T}x{N}x{K}", t, 2 * T * N * K / t / 1e12, "TFLOP/s"))
        tt = timeit(lambda: torch.matmul(a, w.t(), out=out))
        res.append((f"  cublas   {name}", tt, 2 * T * N * K / tt / 1e12, "TFLOP/s"))
        for bn in (64, 128, 256):
            os.environ["APLA_GEMM_BN"] = str(bn)
            t = timeit(lambda: ops.gemm_bias(a, w, bias, out=out))
            res.append((f"  BN={bn}", t, 2 * T * N * K / t / 1e12, "TFLOP/s"))
        del os.environ["APLA_GEMM_BN"]
    w = (torch.randn(4 * D, D, device=dev) * 0.02).bfloat16()
    h = torch.empty(T, 4 * D, device=dev, dtype=torch.bfloat16); g = torch.empty_like(h)
    t = timeit(lambda: ops.gemm_bias_gelu(x, w, torch.zeros(4 * D, device=dev), h=h, g=g))
    res.append(("gemm_bias_gelu fc1", t, 2 * T * 4 * D * D / t / 1e12, "TFLOP/s"))
    dsave = torch.empty(T, 4 * D, device=dev, dtype=torch.float16)
    t = timeit(lambda: ops.gemm_bias_gelu_dgelu(x, w, torch.zeros(4 * D, device=dev), d=dsave, g=g))
    res.append(("gemm_bias_gelu_dgelu fc1", t, 2 * T * 4 * D * D / t / 1e12, "TFLOP/s"))
    w2 = (torch.randn(D, 4 * D, device=dev) * 0.02).bfloat16()
    resid = torch.randn(T, D, device=dev)
    t = timeit(lambda: ops.gemm_bias_ls_residual(x4, w2, None, None, resid, out=resid))
    res.append(("gemm_ls_residual fc2", t, 2 * T * 4 * D * D / t / 1e12, "TFLOP/s"))
    wp = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
    t = timeit(lambda: ops.gemm_bias_ls_residual(x, wp, None, None, resid, out=resid))
    res.append(("gemm_ls_residual proj", t, 2 * T * D * D / t / 1e12, "TFLOP/s"))
    gam = torch.ones(D, device=dev); bb = torch.zeros(D, device=dev)
    t = timeit(lambda: ops.gemm_bias_ls_accumulate(x4, w2, bb, gam, resid))
    res.append(("gemm_ls_accumulate fc2 (TMA reduce-add)", t, 2 * T * 4 * D * D / t / 1e12, "TFLOP/s"))
    t = timeit(lambda: ops.gemm_bias_ls_accumulate(x, wp, bb, gam, resid))
    res.append(("gemm_ls_accumulate proj (TMA reduce-add)", t, 2 * T * D * D / t / 1e12, "TFLOP/s"))
    w2t = w2.t().contiguous()
    dh = torch.empty(T, 4 * D, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.gemm_dgrad_gelu_bwd(x, w2t, h, out=dh))
    res.append(("gemm_dgrad_gelu_bwd fc2", t, 2 * T * 4 * D * D / t / 1e12, "TFLOP/s"))
    t = timeit(lambda: ops.gemm_dgrad_mul(x, w2t, dsave, out=dh))
    res.append(("gemm_dgrad_mul fc2", t, 2 * T * 4 * D * D / t / 1e12, "TFLOP/s"))
    # LN
    xf = torch.randn(T, D, device=dev); wln = torch.ones(D, device=dev); bln = torch.zeros(D, device=dev)
    y = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.layernorm_fwd(xf, wln, bln, 1e-6, out=y))
    res.append(("layernorm_fwd", t, T * D * 6 / t / 1e9, "GB/s"))
    dres = torch.randn(T, D, device=dev); dxb = torch.empty_like(y)
    t = timeit(lambda: ops.layernorm_bwd(y, xf, wln, 1e-6, dres=dres, dx=dres, dxb=dxb))
    res.append(("layernorm_bwd(+resid,+bf16)", t, T * D * (2 + 4 + 4 + 4 + 2) / t / 1e9, "GB/s"))
    # attention
    B, N, H = 64, 257, 12
    qkv = torch.randn(T, 3 * D, device=dev).bfloat16()
    out = torch.empty(T, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(T, H, device=dev)
    t = timeit(lambda: ops.attn_fwd(qkv, H, 0.125, B, N, out=out, lse=lse))
    res.append(("attn_fwd N=257", t, 4 * N * D * T / t / 1e12, "TFLOP/s"))
    dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
    t = timeit(lambda: ops.attn_bwd(qkv, out, y, lse, H, 0.125, B, N, dqkv=dqkv, delta=delta))
    res.append(("attn_bwd N=257", t, 2.5 * 4 * N * D * T / t / 1e12, "TFLOP/s"))
    B, N = 8, 1370
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    t = timeit(lambda: ops.attn_fwd(qkv, H, 0.125, B, N))
    res.append(("attn_fwd N=1370", t, 4 * N * D * B * N / t / 1e12, "TFLOP/s"))
    # wgrad
    for r in (8, 128):
        npad = (r + 63) // 64 * 64
        sub = torch.randn(T, npad, device=dev).bfloat16(); dw = torch.zeros(r, D, device=dev)
        t = timeit(lambda: ops.proj_wgrad(sub, x, dw, r))
        res.append((f"proj_wgrad r={r}", t, 2 * T * D * r / t / 1e12, "TFLOP/s"))
    rowmap = torch.arange(D, dtype=torch.int32, device=dev); dw = torch.zeros(D, D, device=dev)
    t = timeit(lambda: ops.proj_wgrad(x, x, dw, D, rowmap=rowmap))
    res.append(("proj_wgrad r=768", t, 2 * T * D * D / t / 1e12, "TFLOP/s"))
    for name, t, v, u in res:
        print(f"{name:40s} {t * 1e6:10.1f} us  {v:10.1f} {u}")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/kernels.json", "w") as f:
        json.dump([dict(name=n, us=t * 1e6, value=v, unit=u) for n, t, v, u in res], f, indent=1)


if __name__ == "__main__":
    main()
