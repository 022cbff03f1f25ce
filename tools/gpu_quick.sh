#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -q -x --timeout 300 2>&1 | tail -12
for v in 1 0 1 0; do
B2S_PIPELINE=$v timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench_$v.json 2> gpurun_out/quick_bench.err; echo "pipeline=$v rc=$?"
python tools/show_bench.py gpurun_out/quick_bench_$v.json 2>&1 | head -1 | cut -c1-200; grep -v Warning gpurun_out/quick_bench.err | tail -3
done
