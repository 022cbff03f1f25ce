#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_input_pipeline.py tests/test_gpu_coords.py -q -x --timeout 120 2>&1 | tail -12
