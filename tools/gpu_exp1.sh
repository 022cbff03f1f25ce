#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/umma_probe tools/umma_probe.cu 2>&1 | tail -3
echo "=== probe ==="; timeout 60 ./tools/umma_probe | tail -3 | tee gpurun_out/probe.log
for R in 0 1; do
  echo "=== model parity, B2S_EXP_ROUND_A=$R ==="
  rm -f gpurun_out/parity_round$R.jsonl
  B2S_EXP_ROUND_A=$R B2S_PARITY_REPORT=$PWD/gpurun_out/parity_round$R.jsonl timeout -k 10 600 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "AssertionError|passed|failed" | head -20
  cat gpurun_out/parity_round$R.jsonl
  B2S_EXP_ROUND_A=$R timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
done
