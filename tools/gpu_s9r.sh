#!/bin/bash
# MSENet50 (BASELINE.json configs[2]) through the captured-graph training step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 500 python bench.py --model SENet50 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_senet50.log 2> gpurun_out/bench_senet50.err
echo "rc=$?"; tail -3 gpurun_out/bench_senet50.err | cut -c1-300
tail -1 gpurun_out/bench_senet50.log > gpurun_out/bench_senet50.json; python tools/show_bench.py gpurun_out/bench_senet50.json 2>/dev/null | head -12
