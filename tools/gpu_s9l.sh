#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SWEEP_WG='[{}]' SWEEP_TC='[{"tc_zmode":0},{"tc_zmode":1},{"tc_zmode":2},{"tc_zmode":1}]' timeout -k 10 300 python tools/sweep.py 2>&1 | grep -E "^fwd" | tee gpurun_out/sweep_z.log
