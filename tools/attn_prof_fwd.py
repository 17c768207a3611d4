"""Runs the forward attention kernel once at the C2 shape (for instrumented builds)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops
B, N, H = 64, 257, 12
D, T = H * 64, B * N
qkv = torch.randn(T, 3 * D, device="cuda").bfloat16()
out = torch.empty(T, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(T, H, device="cuda")
for _ in range(int(os.environ.get("REPS", "2"))):
    ops.attn_fwd(qkv, H, 0.125, B, N, out=out, lse=lse)
torch.cuda.synchronize()
