#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -p no:cacheprovider -k "attention" 2>&1 | tail -2
for b in 2 8; do
timeout 300 python bench.py --workload c5 --batch $b --steps 30 --warmup 5 --no-cpu --no-sustained 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('c5 b$b', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['step']['frac'],4), d['clocks']['sm_mhz'])"
done
