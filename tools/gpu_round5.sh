#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for conf in "" "roundup_power2_divisions:8" "roundup_power2_divisions:4,max_split_size_mb:512" "expandable_segments:True"; do
  if [ -z "$conf" ]; then unset PYTORCH_CUDA_ALLOC_CONF; else export PYTORCH_CUDA_ALLOC_CONF="$conf"; fi
  PROFILE=0 timeout -k 10 300 python tools/host_profile.py 2>&1 | grep "no profiler" | tee -a gpurun_out/alloc_conf.log
done
