#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/prof_bwd2_r2a -f python tools/attn_check.py --notest > gpurun_out/r2d_ncu.log 2>&1
tail -5 gpurun_out/r2d_ncu.log
ls -la gpurun_out/*.ncu-rep
