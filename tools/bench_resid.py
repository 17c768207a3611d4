"""Residual GEMM (proj / fc2 shapes of the C2 step) timed back to back over rotating operand sets (nothing L2-hot),
CUDA events around 30 launches.  Run once per setting of the environment switch being compared."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops  # noqa: E402


def main():
    dev = "cuda"
    T, D = 16448, 768
    out = {}
    for name, K in (("proj", D), ("fc2", 4 * D)):
        sets = []
        for _ in range(5):
            sets.append((torch.randn(T, K, device=dev).bfloat16(), (torch.randn(D, K, device=dev) * 0.02).bfloat16(),
                         torch.randn(T, D, device=dev), torch.empty(T, D, device=dev)))
        bias = torch.zeros(D, device=dev); gam = torch.ones(D, device=dev)
        def run(n):
            for i in range(n):
                a, w, r, o = sets[i % len(sets)]
                ops.gemm_bias_ls_residual(a, w, bias, gam, r, out=o)
        run(5)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(30); b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 30 * 1e3)
        out[name] = round(best, 2)
    out["env"] = {k: v for k, v in os.environ.items() if k.startswith("APLA_")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
