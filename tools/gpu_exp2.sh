#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for IMPL in 0 1; do
  rm -f gpurun_out/parity_impl$IMPL.jsonl
  B2S_CONV_IMPL=$IMPL B2S_PARITY_REPORT=$PWD/gpurun_out/parity_impl$IMPL.jsonl timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q -p no:cacheprovider -k forward_backward 2>&1 | grep -E "^E  .*Error|passed|failed" | head -20
  echo "--- impl $IMPL"; cat gpurun_out/parity_impl$IMPL.jsonl
done
