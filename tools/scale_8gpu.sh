#!/bin/bash
# 8-GPU job: all-reduce microbenchmark (NVLS on / off), DP correctness, C2 + C3 step with the native and the NCCL reducer,
# C4 and C5 steps
set -u
mkdir -p gpurun_out
N=8
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
APLA_DP_MULTIMEM=1 run 200 29541 tools/bench_allreduce.py > gpurun_out/r2i_allreduce_8gpu_mc.json 2> gpurun_out/r2i_allreduce_8gpu_mc.err; echo "ar mc rc=$?"
grep "^{" gpurun_out/r2i_allreduce_8gpu_mc.json
APLA_DP_MULTIMEM=0 run 200 29542 tools/bench_allreduce.py > gpurun_out/r2i_allreduce_8gpu_nomc.json 2> /dev/null; echo "ar nomc rc=$?"
grep "^{" gpurun_out/r2i_allreduce_8gpu_nomc.json
run 300 29533 tools/dp_check.py > gpurun_out/r2i_dp_check_8gpu.json 2> gpurun_out/r2i_dp_check_8gpu.err; echo "dp_check rc=$?"
grep "^{" gpurun_out/r2i_dp_check_8gpu.json | cut -c1-150
for mode in native nccl; do
APLA_DP_ALLREDUCE=$mode run 300 2953$([ $mode = native ] && echo 4 || echo 5) bench.py --gpus $N --steps 50 --warmup 5 --no-sustained > gpurun_out/r2i_bench_8gpu_$mode.json 2> gpurun_out/r2i_bench_8gpu_$mode.err; echo "bench $mode rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2i_bench_8gpu_$mode.json") if l.startswith("{")][-1])
    print("$mode c2", round(d["value"],1), round(d["ms_per_step"],4), d.get("dp_check"), "launches/step", d["gpu_launches_per_step"], d["clocks"])
    c3=d.get("c3"); print("$mode c3", round(c3["value"],1), round(c3["ms_per_step"],4), c3["params_identical_across_ranks"])
except Exception as e:
    print("ERR", e)
PY
done
APLA_DP_ALLREDUCE=native APLA_DP_MULTIMEM=0 run 300 29536 bench.py --gpus $N --steps 50 --warmup 5 --no-sustained > gpurun_out/r2i_bench_8gpu_native_nomc.json 2> /dev/null; echo "bench native nomc rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2i_bench_8gpu_native_nomc.json") if l.startswith("{")][-1])
    print("native-nomc c2", round(d["value"],1), round(d["ms_per_step"],4)); c3=d.get("c3"); print("native-nomc c3", round(c3["value"],1), round(c3["ms_per_step"],4))
except Exception as e:
    print("ERR", e)
PY
run 400 29537 bench.py --gpus $N --workload c4 --steps 5 --warmup 3 > gpurun_out/r2i_bench_c4_8gpu.json 2> gpurun_out/r2i_bench_c4_8gpu.err; echo "c4 rc=$?"
cut -c1-330 gpurun_out/r2i_bench_c4_8gpu.json
run 300 29538 bench.py --gpus $N --workload c5 --steps 30 --warmup 5 --no-sustained > gpurun_out/r2i_bench_c5_8gpu.json 2> gpurun_out/r2i_bench_c5_8gpu.err; echo "c5 rc=$?"
cut -c1-330 gpurun_out/r2i_bench_c5_8gpu.json
run 300 29539 bench.py --gpus $N --workload c5 --batch 8 --steps 30 --warmup 5 --no-sustained > gpurun_out/r2i_bench_c5b8_8gpu.json 2> /dev/null; echo "c5 b8 rc=$?"
cut -c1-330 gpurun_out/r2i_bench_c5b8_8gpu.json
