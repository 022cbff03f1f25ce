#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2; do echo "B2S_POOL_BATCH=$v"; B2S_POOL_BATCH=$v timeout 120 python tools/pool_bench.py 2>&1 | grep -v Warn | tail -1; done
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_f3.py -q -x --timeout 120 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench_pool.json 2> gpurun_out/quick_bench.err; echo "rc=$?"
python tools/show_bench.py gpurun_out/quick_bench_pool.json 2>&1 | head -1 | cut -c1-200; grep -v Warning gpurun_out/quick_bench.err | tail -3
