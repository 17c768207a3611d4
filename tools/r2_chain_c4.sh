#!/bin/bash
set -u
mkdir -p gpurun_out
for c in 0 1 0 1; do
  APLA_BLOCK_CHAIN=$c timeout 900 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/c4ab_$c.json 2> gpurun_out/c4ab_$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/c4ab_$c.json").read().strip().splitlines()[-1])
print("chain=$c c4", round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", d["clocks"]["sm_mhz"], "launches", d.get("gpu_launches_per_step"))
PY
done
