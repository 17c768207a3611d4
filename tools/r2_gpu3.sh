#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/attn_check.py 2>&1 | grep -v "Warning\|run_backward" > gpurun_out/r2c_attn_v2.txt; echo "rc=$?" >> gpurun_out/r2c_attn_v2.txt
grep "FAIL\|attn_\|Error\|rc=" gpurun_out/r2c_attn_v2.txt
