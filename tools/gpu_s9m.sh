#!/bin/bash
# final evidence of the session: ncu launch list of the bench command (graph replays only: the eager statistics /
# planning passes are skipped, not profiled) + ncu --set full of the hot kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== launch list ==="
( time timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --launch-skip 4300 -c 1400 --csv --log-file gpurun_out/r01_s9_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01_s9_ncu_bench_final.log 2>&1 ) 2>&1 | grep real
wc -l gpurun_out/r01_s9_launches_final.csv
echo "=== full set on the hot kernels ==="
( time REPS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'gather_gemm_tc_kernel|gather_gemm_tc2_kernel|wgrad_small_tc_kernel|wgrad_group_kernel' -c 9 -f -o gpurun_out/r01_s9_hot_final python tools/ncu_target.py > gpurun_out/r01_s9_ncu_hot_final.log 2>&1 ) 2>&1 | grep real
ls -la gpurun_out/r01_s9_hot_final.ncu-rep
