#!/bin/bash
# ncu evidence: launch list of the bench command + one --set full capture of the hot kernels (1 GPU)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01}
echo "=== launch list (bench, graph replays) ==="
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-200; wc -l gpurun_out/${TAG}_launches.csv
echo "=== full set on the hot kernels ==="
REPS=2 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:'gather_gemm_tc_kernel|wgrad_group_kernel|wgrad_small_tc_kernel|kernel_map_kernel' -s 14 -c 14 -f -o gpurun_out/${TAG}_hot python tools/ncu_target.py > gpurun_out/${TAG}_ncu_hot.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_hot.log; ls -la gpurun_out/${TAG}_hot.ncu-rep
