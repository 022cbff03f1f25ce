#!/bin/bash
mkdir -p gpurun_out
for u in 1 4; do echo "B2S_PW_UNROLL=$u"; B2S_PW_UNROLL=$u timeout 200 python tools/bn_bench.py 2>&1 | grep -v Warn | tail -6; done
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x --timeout 120 2>&1 | tail -3
