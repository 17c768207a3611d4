#!/bin/bash
# A/B of the side-stream weight gradient: engine parity tests, then C2 and C3 with APLA_SIDE_WGRAD=0 / 1
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
for sw in 0 1; do
  for wl in c2 c3; do
    APLA_SIDE_WGRAD=$sw timeout 600 python bench.py --workload $wl --steps 50 --warmup 5 --no-c3 --no-cpu --no-sustained > gpurun_out/side_${wl}_$sw.json 2> gpurun_out/side_${wl}_$sw.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/side_${wl}_$sw.json').read().strip().splitlines()[-1])
print('side=$sw $wl ms/step', round(d['ms_per_step'],4), 'img/s', round(d['value'],1), 'loss', d.get('loss'), 'launches/step', d.get('gpu_launches_per_step'), d['clocks']['sm_mhz'])
PY
  done
done
