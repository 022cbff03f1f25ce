#!/bin/bash
# ncu launch list of the bench command (graph nodes profiled individually)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 8000 --csv --log-file gpurun_out/r01_s9_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_s9_ncu_bench.log 2>&1
tail -1 gpurun_out/r01_s9_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r01_s9_launches.csv
