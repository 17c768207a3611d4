"""Attention kernels on the B200: parity against an fp32 torch restatement on several shapes, then CUDA-event timing
(L2 flushed) at the C2 shape (64 x 257 tokens, 12 heads).  Development tool; tests/test_kernels_gpu.py is the gate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops  # noqa: E402

dev = "cuda"


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def ref_attn(qkv, seqlens, H, scale):
    outs, o, D = [], 0, H * 64
    for n in seqlens:
        t = qkv[o:o + n].reshape(n, 3, H, 64).permute(1, 2, 0, 3)
        a = ((t[0] @ t[1].transpose(-2, -1)) * scale).softmax(-1)
        outs.append((a @ t[2]).transpose(0, 1).reshape(n, D))
        o += n
    return torch.cat(outs, 0)


def check(seqlens, H, dense, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    D, T = H * 64, sum(seqlens)
    qkv = torch.randn(T, 3 * D, device=dev, generator=g).bfloat16()
    cu = None if dense else torch.tensor([0] + list(torch.tensor(seqlens).cumsum(0)), dtype=torch.int32, device=dev)
    out, lse = ops.attn_fwd(qkv, H, 0.125, len(seqlens), max(seqlens), cu_seqlens=cu)
    q32 = qkv.float().requires_grad_(True)
    ref = ref_attn(q32, seqlens, H, 0.125)
    dout = torch.randn(T, D, device=dev, generator=g).bfloat16()
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, dout, lse, H, 0.125, len(seqlens), max(seqlens), cu_seqlens=cu)
    torch.cuda.synchronize()
    e = [rel(out.float(), ref.detach())] + [rel(dqkv[:, i * D:(i + 1) * D].float(), q32.grad[:, i * D:(i + 1) * D])
                                            for i in range(3)]
    ok = e[0] < 6e-3 and max(e[1:]) < 1.2e-2
    print(f"{'dense' if dense else 'varlen'} n={seqlens[:4]}{'...' if len(seqlens) > 4 else ''} x{len(seqlens)} H={H}: "
          f"out {e[0]:.2e} dq {e[1]:.2e} dk {e[2]:.2e} dv {e[3]:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def timeit(fn, iters=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
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
    return ts[len(ts) // 2] * 1e3


def main():
    ok = True
    if "--notest" not in sys.argv:
        ok &= check([257] * 2, 12, True)
        ok &= check([197] * 3, 6, True)
        ok &= check([50] * 4, 16, True)
        ok &= check([64] * 2, 2, True)
        ok &= check([17] * 5, 2, True)
        ok &= check([272] * 2, 2, True)
        for n in (65, 128, 129, 193, 256, 63):          # odd-token edge cases of the second-generation backward
            ok &= check([n] * 3, 2, True, seed=n)
        ok &= check([257] * 300, 3, True, seed=9)           # several groups per CTA
        ok &= check([257, 65, 256, 1, 129, 64, 257, 257, 200, 193], 2, False, seed=4)
        ok &= check([257] * 6 + [50] * 90, 16, False, seed=6)   # multi-crop: many single-chunk groups per CTA
        ok &= check([257] * 30, 12, True, seed=3)
        ok &= check([257, 50, 257, 50, 50, 3, 130], 4, False)
        ok &= check([1370], 12, True)
    B, N, H = 64, 257, 12
    D, T = H * 64, B * N
    qkv = torch.randn(T, 3 * D, device=dev).bfloat16()
    dout = torch.randn(T, D, device=dev).bfloat16()
    out = torch.empty(T, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(T, H, device=dev)
    dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
    t = timeit(lambda: ops.attn_fwd(qkv, H, 0.125, B, N, out=out, lse=lse))
    print(f"attn_fwd 64x257x12: {t:.1f} us  {4 * N * D * T / t / 1e6:.0f} TFLOP/s")
    t = timeit(lambda: ops.attn_bwd(qkv, out, dout, lse, H, 0.125, B, N, dqkv=dqkv, delta=delta))
    print(f"attn_bwd 64x257x12: {t:.1f} us  {10 * N * D * T / t / 1e6:.0f} TFLOP/s")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
