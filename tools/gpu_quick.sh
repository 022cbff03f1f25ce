#!/bin/bash
mkdir -p gpurun_out
for z in 0 1; do
echo "B2S_TC_ZST=$z"
B2S_TC_ZST=$z TA_MODES=0 timeout 120 python tools/ta_bench.py 2>&1 | grep -v Warning | tail -4
done
