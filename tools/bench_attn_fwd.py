"""Forward attention at the C2 shape (64 x 257 tokens, 12 heads) timed over rotating inputs: median of 9 x 20 launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops  # noqa: E402

B, N, H = 64, 257, 12
D, T = H * 64, B * N
sets = [(torch.randn(T, 3 * D, device="cuda").bfloat16(), torch.empty(T, D, device="cuda", dtype=torch.bfloat16),
         torch.empty(T, H, device="cuda")) for _ in range(4)]


def run(n):
    for i in range(n):
        q, o, l = sets[i % 4]
        ops.attn_fwd(q, H, 0.125, B, N, out=o, lse=l)


run(8)
torch.cuda.synchronize()
ts = []
for _ in range(9):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(20); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 20 * 1e3)
ts.sort()
print(f"attn_fwd {B}x{N}x{H}: median {ts[4]:.2f} us  min {ts[0]:.2f} us   ({os.environ.get('TAG', '')})")
