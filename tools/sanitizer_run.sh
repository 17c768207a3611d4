#!/bin/bash
# compute-sanitizer evidence for the tcgen05 / mbarrier kernels: memcheck on the kernel-level parity tests,
# racecheck + synccheck on the attention and GEMM cases with small shapes
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
  -k "not 1370 and not 30-257" > gpurun_out/r2h_memcheck_kernels.txt 2>&1
tail -6 gpurun_out/r2h_memcheck_kernels.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
  -k "attention_dense and (2-257-12 or 4-50-16 or 2-64-2) or attention_varlen and 4-seqlens0" > gpurun_out/r2h_racecheck_attn.txt 2>&1
tail -6 gpurun_out/r2h_racecheck_attn.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider \
  -k "attention_dense and (2-257-12 or 4-50-16) or gemm" > gpurun_out/r2h_synccheck.txt 2>&1
tail -6 gpurun_out/r2h_synccheck.txt
