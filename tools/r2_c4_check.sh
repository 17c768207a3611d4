#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_block_gpu.py tests/test_ssl_gpu.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --workload c4 --steps 10 --warmup 4 > gpurun_out/r2n_bench_c4.json 2> gpurun_out/r2n_bench_c4.err; echo "c4 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench_c4.json').read().strip().split('\n')[-1])
print('c4', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches_per_step'], d['clocks'])"
