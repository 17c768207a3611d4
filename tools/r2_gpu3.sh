#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python - > gpurun_out/r2c_attn_prof.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
from apla_b200 import ops
B, N, H = 64, 257, 12
D, T = H * 64, B * N
qkv = torch.randn(T, 3 * D, device='cuda').bfloat16()
dout = torch.randn(T, D, device='cuda').bfloat16()
out, lse = ops.attn_fwd(qkv, H, 0.125, B, N)
dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
for i in range(2):
    ops.attn_bwd(qkv, out, dout, lse, H, 0.125, B, N, dqkv=dqkv, delta=delta)
    torch.cuda.synchronize()
    print("----")
PY
grep afb gpurun_out/r2c_attn_prof.txt | tail -12
