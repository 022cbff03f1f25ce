#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
PW_CFGS='[{"cr_v4":0},{"cr_v4":1,"cr_cap":1},{"cr_v4":1,"cr_cap":2}]' timeout -k 10 300 python tools/pw_bench.py 2>&1 | tail -12 | tee gpurun_out/pw_bench.log
