#!/bin/bash
mkdir -p gpurun_out
export SKIP_STEM=1
export SWEEP_WG='[{}, {"wg_occ2":0}, {"wg_occ2":0,"wg_lag":1}, {"wg_occ2":0,"wg_lag":3}, {"wg_occ2":0,"wg_nbp":32}, {"wg_occ2":0,"wg_wv":8}, {"wg_occ2":0,"wg_wv":6,"wg_lag":1}, {"wg_occ2":0,"wg_ca":1}]'
export SWEEP_TC='[{}]'
timeout 1200 python tools/sweep.py 2>&1 | grep -E "^wgrad|Error" | cut -c1-420
