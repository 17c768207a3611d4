#!/bin/bash
# final-tree verification on one B200: GPU tests, smoke, default bench (the driver's command)
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2z_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2z_bench_default.json 2> gpurun_out/r2z_bench_default.err; cat gpurun_out/r2z_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; cat gpurun_out/r2z_bench_reference.json
