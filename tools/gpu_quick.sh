#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lines.py -q -x --timeout 300 2>&1 | tail -3
for v in 0 1 2 3; do echo "B2S_LINES_WG=$v"; B2S_LINES_WG=$v ONLY_STEM=1 timeout 600 python tools/conv_bench.py 2>&1 | grep -E "stem|Error|error" | cut -c150-400; done
