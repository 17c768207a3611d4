#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "dp_matches or adamw" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py > gpurun_out/r2g_dp_check.json 2> gpurun_out/r2g_dp_check.err; echo "dp_check rc=$?"
grep "^{" gpurun_out/r2g_dp_check.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2g_bench_2gpu.json") if l.startswith("{")][-1])
    print("c2 2gpu", d["value"], d["ms_per_step"], d.get("dp_check"), "sustained", d["roofline"]["step"].get("sustained_images_per_s"))
    c3=d.get("c3"); print("c3", c3["value"], c3["ms_per_step"], c3["params_identical_across_ranks"])
except Exception as e:
    print("ERR", e)
PY
