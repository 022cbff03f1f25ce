#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SWEEP_WG='[{"wg_wv":4},{"wg_wv":2},{"wg_wv":1},{"wg_wv":3},{"wg_wv":6},{"wg_wv":2,"wg_occ2":0},{"wg_wv":4,"wg_nbp":32,"wg_occ2":0}]' SWEEP_TC='[{}]' timeout -k 10 300 python tools/sweep.py 2>&1 | grep wgrad | tee gpurun_out/sweep_k.log
