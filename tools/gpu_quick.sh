#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_model.py tests/test_gpu_conv.py -q -x --timeout 300 2>&1 | tail -3
for v in 1 0 1 0; do
B2S_PREBUILT_IMAGES=$v timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench_$v.json 2> gpurun_out/quick_bench.err; echo "prebuilt images=$v rc=$?"
python tools/show_bench.py gpurun_out/quick_bench_$v.json 2>&1 | head -1 | cut -c1-200; tail -2 gpurun_out/quick_bench.err
done
