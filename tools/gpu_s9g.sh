#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
SWEEP_WG='[{}]' SWEEP_TC='[{}]' timeout -k 10 300 python tools/sweep.py 2>&1 | tail -6 | tee gpurun_out/sweep_g.log
