#!/bin/bash
mkdir -p gpurun_out
export B2S_PARITY_REPORT=gpurun_out/parity_report_r2.jsonl
rm -f $B2S_PARITY_REPORT
timeout 1800 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?"; tail -30 gpurun_out/r2e_tests.log
ONLY_STEM=1 timeout 600 python tools/conv_bench.py 2>&1 | grep "stem"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2e_bench.json 2>/dev/null | head -14
