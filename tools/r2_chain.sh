#!/bin/bash
set -u
mkdir -p gpurun_out
for c in 0 1; do
  APLA_BLOCK_CHAIN=$c timeout 600 python tools/bench_module_paths.py --paths fused_all,fused_all_r768 --steps 30 --warmup 5 2>/dev/null | sed "s/^/chain=$c /"
done
timeout 900 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/r2v_c4.json 2> gpurun_out/r2v_c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_c4.json").read().strip().splitlines()[-1])
print("c4", round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", round(d["roofline"]["frac"], 4), d["clocks"]["sm_mhz"], "launches", d.get("gpu_launches_per_step"))
PY
