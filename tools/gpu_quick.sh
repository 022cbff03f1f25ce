#!/bin/bash
mkdir -p gpurun_out
TA_LAYERS=1 timeout 300 python tools/ta_bench.py 2>&1 | grep -v Warning | tail -3
for v in 0 1; do
B2S_TC_TA=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/quick_bench_ta$v.json 2> gpurun_out/quick_bench.err; echo "tc_ta=$v rc=$?"
python tools/show_bench.py gpurun_out/quick_bench_ta$v.json 2>&1 | head -1 | cut -c1-200; grep -v Warning gpurun_out/quick_bench.err | tail -3
done
