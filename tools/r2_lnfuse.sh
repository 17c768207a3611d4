#!/bin/bash
# the residual GEMM + LayerNorm in one launch: kernel test, engine parity, A/B bench (APLA_GEMM_LN_FUSE=0 / 1)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "layernorm_one_launch or residual" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
for f in 0 1; do
  APLA_GEMM_LN_FUSE=$f timeout 600 python bench.py --steps 50 --warmup 5 --no-c3 --no-cpu --no-sustained > gpurun_out/lnfuse_$f.json 2> gpurun_out/lnfuse_$f.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/lnfuse_$f.json').read().strip().splitlines()[-1])
print('fuse=$f ms/step', round(d['ms_per_step'],4), 'img/s', round(d['value'],1), 'loss', d.get('loss'), 'launches/step', d.get('gpu_launches_per_step'), d['clocks']['sm_mhz'])
PY
done
