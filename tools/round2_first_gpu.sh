#!/bin/bash
# First GPU job of round 2: the self-supervised row kernels (csrc/ssl.cu) have never run on hardware.
#   gpurun --timeout 900 -- 'bash tools/round2_first_gpu.sh'
# 1. parity tests with the xfail marks off, 2. compute-sanitizer on the small cases, 3. per-kernel HBM fractions,
# 4. launch list of the kernel bench.  Everything lands in gpurun_out/ssl_r2_*.
set -u
mkdir -p gpurun_out
APLA_B200_SSL_STRICT=1 timeout 600 python -m pytest tests/test_ssl_gpu.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/ssl_r2_pytest.txt
APLA_B200_SSL_STRICT=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_ssl_gpu.py -q -m gpu \
    -k "not 65536 and not 1048576" 2>&1 | tail -40 > gpurun_out/ssl_r2_memcheck.txt
timeout 600 python tools/bench_ssl_kernels.py > gpurun_out/ssl_r2_kernels.jsonl 2> gpurun_out/ssl_r2_kernels.err
APLA_SSL_SPLIT_CE=1 timeout 600 python tools/bench_ssl_kernels.py 2> /dev/null | grep ssl_objective >> gpurun_out/ssl_r2_kernels.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 \
    --csv --log-file gpurun_out/ssl_r2_launches.csv python tools/bench_ssl_kernels.py > /dev/null 2>&1
tail -5 gpurun_out/ssl_r2_pytest.txt; tail -3 gpurun_out/ssl_r2_memcheck.txt; cat gpurun_out/ssl_r2_kernels.jsonl
# 5. the whole C4-shape step (module-level path), small first, then the real shape
timeout 300 python tools/bench_ssl_step.py --arch vit_base --images 8 --K 4096 --steps 3 --warmup 2 > gpurun_out/ssl_r2_step_small.json 2> gpurun_out/ssl_r2_step_small.err
timeout 600 python tools/bench_ssl_step.py --steps 3 --warmup 2 > gpurun_out/ssl_r2_step_c4.json 2> gpurun_out/ssl_r2_step_c4.err
timeout 600 python tools/bench_ssl_step.py --steps 3 --warmup 2 --per-term-losses > gpurun_out/ssl_r2_step_c4_perterm.json 2> gpurun_out/ssl_r2_step_c4_perterm.err
cat gpurun_out/ssl_r2_step_small.json gpurun_out/ssl_r2_step_c4.json gpurun_out/ssl_r2_step_c4_perterm.json; tail -3 gpurun_out/ssl_r2_step_c4.err
