#!/bin/bash
# GPU test suite + smoke; optional short bench (BENCH=1)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.jsonl
export B2S_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider ${PYTEST_ARGS:-} 2>&1 | tail -${TAIL:-40} | tee gpurun_out/pytest_all.log
cat gpurun_out/parity_report.jsonl 2>/dev/null
echo "=== smoke ==="
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
if [ "${BENCH:-0}" = "1" ]; then
  echo "=== bench ==="
  timeout -k 10 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS:-} 2>&1 | tail -1 > gpurun_out/bench_last.log
  python tools/show_bench.py gpurun_out/bench_last.log
fi
