#!/bin/bash
# compute-sanitizer on the kernels added late in round 2: the one-launch residual GEMM + LayerNorm (small shapes), the wide
# column sum, the packed-fp32 GELU epilogue
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
  -k "(layernorm_one_launch and (300-384 or 64-768 or 2000-512 or 1028-768)) or (colsum_wide and (1031 or 515 or 7-256 or 300-64)) or gelu_saved" > gpurun_out/r2z_memcheck_new.txt 2>&1
tail -5 gpurun_out/r2z_memcheck_new.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
  -k "layernorm_one_launch and (300-384 or 64-768)" > gpurun_out/r2z_synccheck_new.txt 2>&1
tail -5 gpurun_out/r2z_synccheck_new.txt
