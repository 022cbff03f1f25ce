#!/bin/bash
# final state: full gpu suite, smoke, bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_final_4.json
python tools/show_bench.py gpurun_out/bench_final_4.json 2>/dev/null | grep -E "value|maxpool"
