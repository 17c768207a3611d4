"""Runs the attention kernels a few times at the C2 shape (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apla_b200 import ops
B, N, H = 64, 257, 12
D, T = H * 64, B * N
qkv = torch.randn(T, 3 * D, device="cuda").bfloat16()
dout = torch.randn(T, D, device="cuda").bfloat16()
out = torch.empty(T, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(T, H, device="cuda")
dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
for _ in range(3):
    ops.attn_fwd(qkv, H, 0.125, B, N, out=out, lse=lse)
    ops.attn_bwd(qkv, out, dout, lse, H, 0.125, B, N, dqkv=dqkv, delta=delta)
torch.cuda.synchronize()
