#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ssl_gpu.py -q -m gpu 2>&1 | tail -2
timeout 600 python tools/c4_glue_probe.py 2>&1 | grep -v Warning | tail -14
timeout 900 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/r2w_c4.json 2> gpurun_out/r2w_c4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_c4.json").read().strip().splitlines()[-1])
print("c4", round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms", round(d["roofline"]["frac"], 4), d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"], 1), "launches", d.get("gpu_launches_per_step"))
PY
