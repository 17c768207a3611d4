#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/bench_module_paths.py --paths fused_block,fused_all,fused_all_r768 --steps 20 > gpurun_out/r2o_module_paths.jsonl 2> gpurun_out/r2o_module_paths.err
cat gpurun_out/r2o_module_paths.jsonl; tail -2 gpurun_out/r2o_module_paths.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 330 --csv --log-file gpurun_out/r2o_launches_modpath.csv python tools/bench_module_paths.py --paths fused_all --steps 2 --warmup 3 > /dev/null 2>&1
