#!/bin/bash
# final-tree verification on two B200s: full GPU suite (incl. the 2-rank data-parallel test), smoke, 2-GPU bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2y_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2y_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y_bench_2gpu.json 2> gpurun_out/r2y_bench_2gpu.err; tail -c 1500 gpurun_out/r2y_bench_2gpu.json
APLA_SIDE_WGRAD=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/dp_check.py > gpurun_out/r2y_dp_check_side.json 2> gpurun_out/r2y_dp_check_side.err; tail -c 600 gpurun_out/r2y_dp_check_side.json
