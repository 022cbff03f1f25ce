#!/bin/bash
# session 9, call F: full gpu suite + smoke + bench after the TF32-twin producers and the column reductions
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
export B2S_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
echo "=== gpu suite ==="
( time timeout -k 10 700 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/pytest_s9f.log
grep twins gpurun_out/parity_report.jsonl
echo "=== smoke ==="
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench ==="
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_s9f.log
python tools/show_bench.py gpurun_out/bench_s9f.log
