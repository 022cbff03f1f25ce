#!/bin/bash
# ncu evidence (1 GPU): launch list of the bench command + one --set full capture of one instance of every stage
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
echo "=== launch list (bench, graph replays) ==="
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200; wc -l gpurun_out/${TAG}_launches.csv
echo "=== full set, one instance per stage ==="
timeout -k 10 1500 ncu --set full --clock-control none --import-source on -f -o gpurun_out/${TAG}_stages python tools/ncu_target.py > gpurun_out/${TAG}_ncu_stages.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_stages.log; ls -la gpurun_out/${TAG}_stages.ncu-rep
ncu -i gpurun_out/${TAG}_stages.ncu-rep --page raw --csv > gpurun_out/${TAG}_stages_raw.csv 2>/dev/null
python tools/ncu_instances.py gpurun_out/${TAG}_stages_raw.csv gpurun_out/ncu_manifest.json gpurun_out/${TAG}_ncu_instances.json MEASURED_PEAKS.json | tee gpurun_out/${TAG}_ncu_instances.txt
