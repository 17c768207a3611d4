#!/bin/bash
# 4-GPU check of the native all-reduce (NVLS path starts at 4 ranks): DP correctness, microbenchmark, C2 + C3 step
set -u
mkdir -p gpurun_out
N=4
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 300 29533 tools/dp_check.py > gpurun_out/r2k_dp_check_4gpu.json 2> gpurun_out/r2k_dp_check_4gpu.err; echo "dp_check rc=$?"
grep "^{" gpurun_out/r2k_dp_check_4gpu.json | cut -c1-120
run 200 29541 tools/bench_allreduce.py > gpurun_out/r2k_allreduce_4gpu.json 2> /dev/null; echo "ar rc=$?"
grep "^{" gpurun_out/r2k_allreduce_4gpu.json
run 300 29534 bench.py --gpus $N --steps 50 --warmup 5 --no-sustained > gpurun_out/r2k_bench_4gpu.json 2> gpurun_out/r2k_bench_4gpu.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2k_bench_4gpu.json") if l.startswith("{")][-1])
    print("c2", round(d["value"],1), round(d["ms_per_step"],4), d.get("dp_check"), "launches/step", d["gpu_launches_per_step"])
    c3=d.get("c3"); print("c3", round(c3["value"],1), round(c3["ms_per_step"],4), c3["params_identical_across_ranks"])
except Exception as e:
    print("ERR", e)
PY
