#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_block_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --workload c3 --steps 50 --warmup 5 --no-c3 --no-cpu > gpurun_out/r2x_c3.json 2> gpurun_out/r2x_c3.err
timeout 900 python bench.py --workload c4 --steps 8 --warmup 3 > gpurun_out/r2x_c4.json 2> gpurun_out/r2x_c4.err
python - <<'PY'
import json
for f in ("r2x_c3", "r2x_c4"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms", d.get("roofline", {}).get("step", {}).get("frac"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-800:])
PY
