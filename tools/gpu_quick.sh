#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/quick_bench.json 2>&1 | head -3; tail -3 gpurun_out/quick_bench.err
