#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for V in 0 10 11; do
  echo "variant $V"
  B2S_TC_M256=0 B2S_TC_VARIANT=$V SWEEP_WG='[]' SWEEP_TC='[{}]' timeout -k 10 300 python tools/sweep.py 2>&1 | grep -E "^fwd" | cut -c1-600
done | tee gpurun_out/sweep_tma.log
