#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/bench_allreduce.py > gpurun_out/r2g_allreduce_2gpu.json 2> gpurun_out/r2g_allreduce_2gpu.err; echo "rc=$?"
grep "^{" gpurun_out/r2g_allreduce_2gpu.json; tail -3 gpurun_out/r2g_allreduce_2gpu.err
APLA_DP_MULTIMEM=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/bench_allreduce.py > gpurun_out/r2g_allreduce_2gpu_nomc.json 2> /dev/null; echo "rc=$?"
grep "^{" gpurun_out/r2g_allreduce_2gpu_nomc.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > gpurun_out/r2g_dp_check.json 2> gpurun_out/r2g_dp_check.err; echo "dp_check rc=$?"
grep "^{" gpurun_out/r2g_dp_check.json | cut -c1-200
APLA_DP_ALLREDUCE=native timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 50 --warmup 5 --no-sustained > gpurun_out/r2g_bench_2gpu_native.json 2> gpurun_out/r2g_bench_2gpu_native.err; echo "bench2 rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2g_bench_2gpu_native.json") if l.startswith("{")][-1])
    print("native c2", round(d["value"],1), round(d["ms_per_step"],4), d.get("dp_check"), "launches/step", d["gpu_launches_per_step"])
    c3=d.get("c3"); print("native c3", round(c3["value"],1), round(c3["ms_per_step"],4), c3["params_identical_across_ranks"])
except Exception as e:
    print("ERR", e)
PY
