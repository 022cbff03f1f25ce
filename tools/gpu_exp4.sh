#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_coords.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15
echo "=== conv bench ==="; timeout -k 10 600 python tools/conv_bench.py 2>&1 | tail -12 | cut -c1-520
