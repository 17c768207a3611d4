#!/bin/bash
# Round-2 GPU job 1: strict SSL parity run (full tracebacks), baseline bench on this box, first C4-step timing.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ssl_gpu.py -q -m gpu --tb=long -p no:cacheprovider > gpurun_out/r2a_ssl_pytest.txt 2>&1
tail -15 gpurun_out/r2a_ssl_pytest.txt
timeout 300 python bench.py > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err
cat gpurun_out/r2a_bench_default.json
timeout 300 python tools/bench_ssl_step.py --arch vit_base --images 8 --K 4096 --steps 3 --warmup 2 > gpurun_out/r2a_step_small.json 2> gpurun_out/r2a_step_small.err
timeout 400 python tools/bench_ssl_step.py --steps 3 --warmup 2 > gpurun_out/r2a_step_c4.json 2> gpurun_out/r2a_step_c4.err
cat gpurun_out/r2a_step_small.json gpurun_out/r2a_step_c4.json; tail -5 gpurun_out/r2a_step_c4.err
timeout 300 python tools/bench_ssl_kernels.py > gpurun_out/r2a_ssl_kernels.jsonl 2> gpurun_out/r2a_ssl_kernels.err
cat gpurun_out/r2a_ssl_kernels.jsonl
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_ssl_gpu.py -q -m gpu -p no:cacheprovider \
    -k "not 65536 and not 1048576 and not dino_head and not ssl_step" > gpurun_out/r2a_memcheck.txt 2>&1
tail -8 gpurun_out/r2a_memcheck.txt
