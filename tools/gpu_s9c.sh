#!/bin/bash
# session 9, call C: conv/ops parity after the kernel changes (L1 gathers, index run-ahead, float4 column reductions),
# knob sweep, short bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gpu tests (conv, ops) ==="
( time timeout -k 10 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/pytest_s9c.log
echo "=== sweep ==="
( time timeout -k 10 500 python tools/sweep.py 2>&1 | tail -60 ) 2>&1 | tee gpurun_out/sweep.log
echo "=== bench ==="
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s9c.log
python tools/show_bench.py gpurun_out/bench_s9c.log
