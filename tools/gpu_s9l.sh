#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
SWEEP_WG='[{}]' SWEEP_TC='[{"tc_m256":0},{"tc_m256":1},{"tc_m256":2},{"tc_m256":3}]' timeout -k 10 300 python tools/sweep.py 2>&1 | grep -E "^fwd" | tee gpurun_out/sweep_l.log
