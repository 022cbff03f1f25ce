#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/noise_check.py 2>&1 | grep -v Warn | tail -3
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest tests/test_gpu_f4.py -q -x --timeout 100 -k checkpoint_with_optimiser 2>&1 | grep -E "passed|failed|^E  +Assert" | head -2; done
