#!/bin/bash
# final tree of round 2: launch list of the default bench command + ncu --set full of one forward block
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c3 --no-sustained > gpurun_out/launches_r2z.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-c3 --no-sustained"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attn_|ln_" -s 175 -c 7 -f -o gpurun_out/prof_fwd_r2z $B > gpurun_out/ncu_fwd_r2z.log 2>&1
ls -la gpurun_out/prof_fwd_r2z.ncu-rep gpurun_out/launches_r2z.csv
