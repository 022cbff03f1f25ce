#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python tools/conv_bench.py 2>&1 | grep -E "s2" | sed -E "s/'n_in.*'fill': [0-9.]+, 'kernel_map_ms': [0-9.]+, 'kernel_map_GBps': [0-9.]+, //" | cut -c1-300
