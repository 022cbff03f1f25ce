#!/bin/bash
# session 8, call A: full gpu suite (state after the wgrad ring changes) + launch-tuning sweep
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "=== gpu suite ==="
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/pytest_s8a.log
echo "=== sweep ==="
timeout -k 10 600 python tools/sweep.py 2>&1 | tail -40 | tee gpurun_out/sweep.log
