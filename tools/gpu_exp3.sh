#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -40
echo "=== bench ==="
timeout -k 10 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 > gpurun_out/bench_last.log
python tools/show_bench.py gpurun_out/bench_last.log || tail -20 gpurun_out/bench_last.log
