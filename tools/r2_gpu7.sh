#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/attn_check.py 2>&1 | grep -v "Warning\|run_backward" > gpurun_out/r2c_attn_v2.txt; echo "rc=$?" >> gpurun_out/r2c_attn_v2.txt
grep "FAIL\|attn_\|Error\|rc=\|multi\|x96" gpurun_out/r2c_attn_v2.txt
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; echo "c4 rc=$?"
grep -v "Warning\|warn\|WeightNorm" gpurun_out/r2f_bench_c4.err | tail -3; cut -c1-2500 gpurun_out/r2f_bench_c4.json
timeout 300 python tools/bench_ssl_kernels.py > gpurun_out/ssl_kernels_r2.jsonl 2>/dev/null; wc -l gpurun_out/ssl_kernels_r2.jsonl
