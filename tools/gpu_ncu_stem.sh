#!/bin/bash
# ncu --set full of the stem instances only (source-level stalls of the x-line kernels)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02_stem}
ONLY_STEM=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:lines -f -o gpurun_out/${TAG} python tools/ncu_target.py > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/${TAG}.ncu-rep
