#!/bin/bash
mkdir -p gpurun_out
export B2S_PARITY_REPORT=gpurun_out/parity_report_r2.jsonl
rm -f $B2S_PARITY_REPORT
timeout 1800 python -m pytest tests -q -m gpu --timeout 300 -x --deselect tests/test_gpu_reference_nets.py > gpurun_out/r2l_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r2l_tests.log
bash tools/gpu_refnets.sh
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2l_bench.json 2>&1 | head -50; tail -3 gpurun_out/r2l_bench.err
timeout 900 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_cfg3.json 2> gpurun_out/r2l_bench_cfg3.err; echo "bench cfg3 rc=$?"
python tools/show_bench.py gpurun_out/r2l_bench_cfg3.json 2>&1 | head -3
