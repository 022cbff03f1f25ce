#!/bin/bash
# round 2, run A: conv parity in both operand modes, then the whole GPU suite with the parity report
mkdir -p gpurun_out
export B2S_PARITY_REPORT=gpurun_out/parity_report_r2.jsonl
rm -f $B2S_PARITY_REPORT
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 120 -x > gpurun_out/r2a_conv.log 2>&1
echo "conv rc=$?"; tail -15 gpurun_out/r2a_conv.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_conv.py > gpurun_out/r2a_all.log 2>&1
echo "all rc=$?"; tail -30 gpurun_out/r2a_all.log
