#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -x -k "se_tail or twins or forward_backward" 2>&1 | tail -15
