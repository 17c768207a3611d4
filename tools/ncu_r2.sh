#!/bin/bash
# ncu --set full captures of round 2 (C2 step, bench.py): a forward block, the new attention backward, and the dominant
# GEMM kernel's three shapes (qkv forward / fc1 dgrad / qkv dgrad) for roofline.traffic
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-sustained"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attn_|ln_" -s 175 -c 7 -f -o gpurun_out/prof_fwd_r2 $B > gpurun_out/ncu_fwd_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd2" -s 12 -c 1 -f -o gpurun_out/prof_attnbwd2_r2 $B > gpurun_out/ncu_attnbwd2_r2.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"gemm2_kernel<256, 0>" -s 50 -c 4 -f -o gpurun_out/prof_gemm0_r2 $B > gpurun_out/ncu_gemm0_r2.log 2>&1
ls -la gpurun_out/prof_*_r2.ncu-rep
