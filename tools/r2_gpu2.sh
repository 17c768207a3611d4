#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ssl_gpu.py -q -m gpu --tb=short -p no:cacheprovider -k "oracle_assembly or ssl_step or no_masked" > gpurun_out/r2b_ssl_pytest.txt 2>&1
tail -30 gpurun_out/r2b_ssl_pytest.txt
