#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
timeout -k 10 300 python bench.py --model SENet50 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_senet50.json
python tools/show_bench.py gpurun_out/bench_senet50.json 2>/dev/null | grep -E "value|se_gate"
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_final_3.json
python tools/show_bench.py gpurun_out/bench_final_3.json 2>/dev/null | grep -E "value|se_gate"
