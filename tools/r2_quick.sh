#!/bin/bash
# quick A/B: GEMM/engine parity tests + the default bench line
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-c3 --no-cpu > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('ms/step', d['ms_per_step'], 'img/s', d['value'], 'sustained', r['step'].get('sustained_ms_per_step'), r['step'].get('sustained_images_per_s'))
print('dominant us', r['us_per_launch'], 'fc1', r['fc1_gelu_kernel'])
PY
