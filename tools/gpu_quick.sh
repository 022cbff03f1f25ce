#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_graph.py tests/test_gpu_f4.py -q -x --timeout 300 2>&1 | tail -5
