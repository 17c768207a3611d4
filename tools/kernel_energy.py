"""Energy per launch of the step's main kernels: each kernel is looped for ~1.2 s over rotating operands while NVML samples
board power (a moving average of about a second: each kernel runs 3 s and the second half of the samples is used);
energy = mean power x time per launch, at whatever clock the power cap leaves that kernel.  The sustained step runs into the board's power cap, so a kernel's
share of the step's ENERGY, not of its time, is what a clock-limited run pays for.  One JSON line per kernel."""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops  # noqa: E402


class Power:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.samples, self.clk, self.run = [], [], False

    def __enter__(self):
        self.samples, self.clk, self.run = [], [], True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()
        return self

    def _loop(self):
        while self.run:
            self.samples.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
            self.clk.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            time.sleep(0.02)

    def __exit__(self, *a):
        self.run = False
        self.t.join()


ONLY = os.environ.get("ONLY", "")


def measure(name, fn, per_step, flops=0.0, mbytes=0.0, seconds=3.0):
    if ONLY and not any(k in name for k in ONLY.split(",")):
        return
    for _ in range(10):
        fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    n = max(50, int(seconds * 1e3 / (a.elapsed_time(b) / 20)))
    with Power() as p:
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    s = p.samples[len(p.samples) // 2:] or p.samples          # NVML power is a ~1 s moving average: use the second half
    c = p.clk[len(p.clk) // 2:] or p.clk
    w = sum(s) / len(s)
    print(json.dumps(dict(kernel=name, us=round(us, 2), watts=round(w, 1), sm_mhz=round(sum(c) / len(c)), mj_per_launch=round(w * us * 1e-3, 3),
                          launches_per_step=per_step, mj_per_step=round(w * us * 1e-3 * per_step, 2),
                          pj_per_flop=round(w * us * 1e-6 / flops * 1e12, 3) if flops else None,
                          pj_per_byte=round(w * us * 1e-6 / (mbytes * 1e6) * 1e12, 1) if mbytes else None)), flush=True)


def main():
    dev = "cuda"
    T, D, H, B, N = 16448, 768, 12, 64, 257
    R = 4
    xs = [torch.randn(T, D, device=dev).bfloat16() for _ in range(R)]
    x4 = [torch.randn(T, 4 * D, device=dev).bfloat16() for _ in range(R)]
    xf = [torch.randn(T, D, device=dev) for _ in range(R)]
    wq = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16(); bq = torch.zeros(3 * D, device=dev)
    w1 = (torch.randn(4 * D, D, device=dev) * 0.02).bfloat16(); b1 = torch.zeros(4 * D, device=dev)
    w2 = (torch.randn(D, 4 * D, device=dev) * 0.02).bfloat16(); wp = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
    bd = torch.zeros(D, device=dev); gam = torch.ones(D, device=dev)
    qkv = [torch.randn(T, 3 * D, device=dev).bfloat16() for _ in range(R)]
    oq = torch.empty(T, 3 * D, device=dev, dtype=torch.bfloat16)
    d16 = torch.empty(T, 4 * D, device=dev, dtype=torch.float16); g16 = torch.empty(T, 4 * D, device=dev, dtype=torch.bfloat16)
    of = torch.empty(T, D, device=dev); ob = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(T, H, device=dev); delta = torch.empty(T, H, device=dev); dqkv = torch.empty(T, 3 * D, device=dev, dtype=torch.bfloat16)
    F = lambda m, n, k: 2.0 * m * n * k                                                           # noqa: E731
    measure("gemm qkv (EPI_BIAS 16448x2304x768)", lambda i: ops.gemm_bias(xs[i % R], wq, bq, out=oq), 11 + 10 + 10 + 4, F(T, 3 * D, D))
    measure("gemm fc1 + GELU + GELU'", lambda i: ops.gemm_bias_gelu_dgelu(xs[i % R], w1, b1, d=d16, g=g16), 11, F(T, 4 * D, D))
    measure("gemm fc2 + LayerScale + residual", lambda i: ops.gemm_bias_ls_residual(x4[i % R], w2, bd, gam, xf[i % R], out=of), 11, F(T, D, 4 * D))
    measure("gemm proj + LayerScale + residual", lambda i: ops.gemm_bias_ls_residual(xs[i % R], wp, bd, gam, xf[i % R], out=of), 11, F(T, D, D))
    hp = torch.randn(T, 4 * D, device=dev).half()
    w2t = w2.t().contiguous(); w1t = w1.t().contiguous(); wpt = wp.t().contiguous()
    dh = torch.empty(T, 4 * D, device=dev, dtype=torch.bfloat16)
    measure("gemm fc2 dgrad x GELU'", lambda i: ops.gemm_dgrad_mul(xs[i % R], w2t, hp, out=dh), 11, F(T, 4 * D, D))
    measure("attention forward", lambda i: ops.attn_fwd(qkv[i % R], H, 0.125, B, N, out=ob, lse=lse), 11, 4.0 * N * D * T)
    ops.attn_fwd(qkv[0], H, 0.125, B, N, out=ob, lse=lse)
    measure("attention backward", lambda i: ops.attn_bwd(qkv[i % R], ob, xs[i % R], lse, H, 0.125, B, N, dqkv=dqkv, delta=delta), 10,
            10.0 * N * D * T)
    measure("layernorm forward", lambda i: ops.layernorm_fwd(xf[i % R], gam, bd, 1e-6, out=ob), 25, mbytes=T * D * 6 / 1e6)
    measure("ls_cast (fp32 -> bf16, same bytes as layernorm forward, no reductions)", lambda i: ops.ls_cast(xf[i % R], gam, out=ob), 0,
            mbytes=T * D * 6 / 1e6)
    dres = torch.randn(T, D, device=dev); dxb = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
    measure("layernorm backward", lambda i: ops.layernorm_bwd(xs[i % R], xf[i % R], gam, 1e-6, dres=dres, dx=dres, dxb=dxb), 24,
            mbytes=T * D * 16 / 1e6)
    ca, cb2 = torch.empty(1 << 28, device=dev), torch.empty(1 << 28, device=dev)           # 1 GiB each
    measure("torch copy_ 1 GiB fp32 (calibration)", lambda i: cb2.copy_(ca), 0, mbytes=2 * (1 << 30) / 1e6)
    big = (torch.randn(8192, 8192, device=dev).bfloat16(), torch.randn(8192, 8192, device=dev).bfloat16(), torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16))
    measure("cuBLAS 8192^3 (calibration)", lambda i: torch.matmul(big[0], big[1].t(), out=big[2]), 0, F(8192, 8192, 8192))


if __name__ == "__main__":
    main()
