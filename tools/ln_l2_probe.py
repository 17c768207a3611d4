"""Probe: how fast is apla_layernorm_fwd when its input is (a) cold in DRAM, (b) just written by the previous kernel,
(c) just written and small enough to sit in L2 for sure?  Decides whether L2 eviction hints on the producing GEMM's stores
are worth building.  One JSON line per case."""
import json
import torch
from apla_b200 import ops


def timed(fn_pre, fn, n=20):
    ts = []
    for _ in range(n):
        fn_pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda:0"
    D = 768
    w = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for rows in (16448, 8224, 4112):
        src = torch.randn(rows, D, device=dev)
        x = torch.empty_like(src)
        y = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
        mb = rows * D * 6 / 1e6
        cold = timed(lambda: (x.copy_(src), flush.zero_()), lambda: ops.layernorm_fwd(x, w, b, 1e-6, out=y))
        warm = timed(lambda: (flush.zero_(), x.copy_(src)), lambda: ops.layernorm_fwd(x, w, b, 1e-6, out=y))
        hot = timed(lambda: (flush.zero_(), x.copy_(src), ops.layernorm_fwd(x, w, b, 1e-6, out=y)),
                    lambda: ops.layernorm_fwd(x, w, b, 1e-6, out=y))
        print(json.dumps(dict(rows=rows, mbytes=round(mb, 1), cold_us=round(cold, 2), after_write_us=round(warm, 2),
                              after_read_us=round(hot, 2), cold_gbs=round(mb / cold * 1e3), warm_gbs=round(mb / warm * 1e3),
                              hot_gbs=round(mb / hot * 1e3))))


if __name__ == "__main__":
    main()
