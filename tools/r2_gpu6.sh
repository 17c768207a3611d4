#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --no-c3 --steps 100 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c2', round(d['value'],1), round(d['ms_per_step'],4), d['clocks']['sm_mhz'], 'sustained', round(d['roofline']['step']['sustained_images_per_s'],1))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 210 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c3 --no-sustained > /dev/null 2>&1
echo done
