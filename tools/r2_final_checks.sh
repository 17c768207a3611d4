#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/parity_report.py > gpurun_out/parity_r2.txt 2>&1; tail -16 gpurun_out/parity_r2.txt
timeout 300 python tools/determinism_check.py > gpurun_out/determinism_r2.txt 2>&1; tail -6 gpurun_out/determinism_r2.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
