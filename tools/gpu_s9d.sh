#!/bin/bash
# session 9, call D: parity after the stem gather remap; default-knob conv timings; column-reduction microbench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gpu tests (conv, ops) ==="
timeout -k 10 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 | tee gpurun_out/pytest_s9d.log
echo "=== sweep (defaults only) ==="
SWEEP_WG='[{}]' SWEEP_TC='[{}]' timeout -k 10 300 python tools/sweep.py 2>&1 | tail -6 | tee gpurun_out/sweep_d.log
echo "=== pw bench ==="
timeout -k 10 300 python tools/pw_bench.py 2>&1 | tail -12 | tee gpurun_out/pw_bench.log
