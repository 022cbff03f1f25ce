#!/bin/bash
# session 9, call A: gpu suite + launch-tuning sweep + short bench (state after the wgrad ring changes)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== gpu suite ==="
( time timeout -k 10 700 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/pytest_s9a.log
echo "=== sweep ==="
( time timeout -k 10 400 python tools/sweep.py 2>&1 | tail -60 ) 2>&1 | tee gpurun_out/sweep.log
echo "=== bench ==="
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s9a.log
python tools/show_bench.py gpurun_out/bench_s9a.log
