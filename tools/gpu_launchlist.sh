#!/bin/bash
# launch list only (1 GPU): the timed graph replays of the default bench command under ncu's duration-only pass
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02b}
B2S_NCU_RANGE=1 timeout -k 10 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200; wc -l gpurun_out/${TAG}_launches.csv
python tools/ncu_launch_summary.py gpurun_out/${TAG}_launches.csv 2 > gpurun_out/${TAG}_launches_summary.txt; head -60 gpurun_out/${TAG}_launches_summary.txt
