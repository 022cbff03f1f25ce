#!/bin/bash
# ncu --set full with source counters on the stem kernels and the L1 64->64 kernels (one launch each)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REPS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:'gather_gemm_tc_kernel|wgrad_small_tc_kernel|wgrad_group_kernel' -c 5 -f -o gpurun_out/r01_s9_hot python tools/ncu_target.py > gpurun_out/r01_s9_ncu_hot.log 2>&1
tail -3 gpurun_out/r01_s9_ncu_hot.log; ls -la gpurun_out/r01_s9_hot.ncu-rep
