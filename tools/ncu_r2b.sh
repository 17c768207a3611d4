#!/bin/bash
set -u
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-sustained"
timeout 500 ncu --set full --clock-control none -k regex:gemm2_kernel -s 52 -c 8 -f -o gpurun_out/prof_gemm_bwd_r2 $B > gpurun_out/ncu_gemm_r2.log 2>&1
ls -la gpurun_out/prof_gemm_bwd_r2.ncu-rep
