#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_f3.py tests/test_gpu_f4.py tests/test_gpu_conv.py tests/test_gpu_graph.py -q -m gpu --timeout 300 > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?"; tail -40 gpurun_out/r2d_tests.log
PRECISE=1 timeout 600 python tools/conv_bench.py 2>&1 | grep "stem"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2d_bench.json | head -14
