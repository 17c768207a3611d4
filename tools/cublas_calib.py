"""Calibration only (not on any product path): what cuBLAS reaches on the step's GEMM shapes, to judge how much
headroom the hand-written 2-CTA kernel's main loop has left.  Prints one JSON line per shape."""
import json
import torch

def timeit(f, n=30):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

def main():
    torch.manual_seed(0)
    dev = "cuda:0"
    M = 16448
    for name, N, K in (("qkv", 2304, 768), ("fc1", 3072, 768), ("fc2", 768, 3072), ("proj", 768, 768), ("big", 8192, 8192)):
        m = M if name != "big" else 8192
        # rotate through enough operand sets that inputs do not sit in L2 between launches
        sets = [(torch.randn(m, K, device=dev, dtype=torch.bfloat16), torch.randn(N, K, device=dev, dtype=torch.bfloat16),
                 torch.empty(m, N, device=dev, dtype=torch.bfloat16)) for _ in range(6)]
        i = [0]
        def f():
            a, b, c = sets[i[0] % len(sets)]
            i[0] += 1
            torch.matmul(a, b.t(), out=c)
        us = timeit(f)
        print(json.dumps(dict(shape=name, M=m, N=N, K=K, us=round(us, 2), tflops=round(2.0 * m * N * K / us / 1e6, 1))))

if __name__ == "__main__":
    main()
