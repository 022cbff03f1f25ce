#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_graph.py -q -x --timeout 200 2>&1 | tail -12
