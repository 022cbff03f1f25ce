#!/bin/bash
for i in 1 2 3 4 5 6; do timeout 900 python -m pytest tests/test_gpu_f4.py tests/test_gpu_f3.py -q --timeout 300 2>&1 | grep -E "passed|failed|grad of|buffer |output|^E  " | head -6; done
