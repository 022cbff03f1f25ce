#!/bin/bash
mkdir -p gpurun_out
export B2S_PARITY_REPORT=gpurun_out/parity_report_r2.jsonl
rm -f $B2S_PARITY_REPORT
timeout 1800 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench_cfg3.json 2> gpurun_out/quick_bench.err; echo "cfg3 rc=$?"
python tools/show_bench.py gpurun_out/quick_bench_cfg3.json 2>&1 | head -1 | cut -c1-200; tail -2 gpurun_out/quick_bench.err
