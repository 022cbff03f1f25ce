#!/bin/bash
# final state: full gpu suite, smoke, two bench runs (run-to-run spread of the last loss)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do
  timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_final_$i.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_final_$i.json')); print({k:d[k] for k in ('value','ms_per_step','last_loss')}, d['e2e']['value'])"
done
