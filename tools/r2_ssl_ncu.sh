#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"softmax_center|soft_ce|colsum_f32|ema_kernel|weightnorm|sk_" -c 120 \
    --csv --log-file gpurun_out/r2p_ssl_dram.csv python tools/bench_ssl_kernels.py > /dev/null 2>&1
wc -l gpurun_out/r2p_ssl_dram.csv
