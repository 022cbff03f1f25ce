#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s9n.log
python tools/show_bench.py gpurun_out/bench_s9n.log | head -3
