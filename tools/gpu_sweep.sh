#!/bin/bash
mkdir -p gpurun_out
export SKIP_STEM=1
export SWEEP_WG='[{}, {"wg_occ2":1}, {"wg_lag":2}, {"wg_lag":3}, {"wg_nbp":32}, {"wg_nbp":16}, {"wg_wv":8}, {"wg_wv":4}, {"wg_occ2":1,"wg_lag":2}, {"wg_ca":1}]'
export SWEEP_TC='[{}, {"tc_m256":3}, {"tc_m256":0}, {"tc_rot":1}, {"tc_ca":1}, {"tc_occ1":1}]'
timeout 600 python tools/sweep.py 2>&1 | grep -E "^wgrad|^fwd|^conv|^tc|Error" | cut -c1-420
