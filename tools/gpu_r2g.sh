#!/bin/bash
# round 2, session 2: x-line stem kernels -- parity first, then per-layer timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lines.py -x -q --timeout 300 > gpurun_out/r2g_lines.log 2>&1
echo "lines tests rc=$?"; tail -25 gpurun_out/r2g_lines.log
for v in 0 3 4 5; do echo "B2S_LINES_FWD=$v"; B2S_LINES_FWD=$v ONLY_STEM=1 timeout 600 python tools/conv_bench.py 2>&1 | grep -E "stem|Error|error" | head -3; done
