#!/bin/bash
set -u
cd "$(dirname "$0")/.."
B2S_TC_VARIANT=10 timeout -k 10 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "forward or backward or identity" 2>&1 | tail -8
for V in 0 10 11; do
  echo "=== variant $V ==="
  B2S_TC_VARIANT=$V timeout -k 10 300 python tools/conv_bench.py 2>&1 | grep -E "layer" | grep -v "stem\|pool" | sed -E "s/'n_in.*'fill': [0-9.]+, 'kernel_map_ms': [0-9.]+, 'kernel_map_GBps': [0-9.]+, //" | cut -c1-260
done
