#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1100 --csv --log-file gpurun_out/r2m_launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 > gpurun_out/r2m_c4.log 2>&1
tail -2 gpurun_out/r2m_c4.log; wc -l gpurun_out/r2m_launches_c4.csv
