#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2; do for nar in 1 0; do
  APLA_COLSUM_NARROW=$nar timeout 600 python bench.py --workload c3 --steps 60 --warmup 5 --no-c3 --no-cpu --no-sustained > gpurun_out/c3ab_$nar.json 2> gpurun_out/c3ab_$nar.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c3ab_$nar.json').read().strip().splitlines()[-1])
print('narrow=$nar', round(d['ms_per_step'],4), 'ms', round(d['value'],1), 'img/s', d['clocks']['sm_mhz'])
PY
done; done
