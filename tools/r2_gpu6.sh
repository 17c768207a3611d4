#!/bin/bash
set -u
mkdir -p gpurun_out
for r in 1 2; do
for v in 0 1; do
APLA_ATTN_BWD_V1=$v timeout 300 python bench.py --no-cpu --steps 100 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('v1=$v', round(d['value'],1), round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
done
APLA_ATTN_BWD_V1=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 210 --csv --log-file gpurun_out/r2e_launches_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2e_launches_v2.csv 2>/dev/null | head -14
