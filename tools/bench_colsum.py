"""Column-sum (bias gradient) kernel on the C3 / C4 shapes: microseconds and GB/s, L2 flushed between launches."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops  # noqa: E402


def main():
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for T, n in ((16448, 768), (58496, 1024), (16448, 64)):
        dy = torch.randn(T, n, device=dev).bfloat16()
        db = torch.zeros(n, device=dev)
        ts = []
        for _ in range(12):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.colsum(dy, db, n); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        print(json.dumps(dict(rows=T, n=n, us=round(us, 2), gbs=round(T * n * 2 / us / 1e3, 1))))


if __name__ == "__main__":
    main()
