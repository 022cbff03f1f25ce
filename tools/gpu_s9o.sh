#!/bin/bash
# end-of-session check, as the driver runs it: default bench (cpu baseline included), reference arm
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout -k 10 600 python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err ) 2>&1 | grep real
tail -1 gpurun_out/bench_default.log > gpurun_out/bench_default.json; python tools/show_bench.py gpurun_out/bench_default.json 2>/dev/null | head -5
( time timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1 ) 2>&1 | grep real
tail -1 gpurun_out/bench_reference.log | cut -c1-400
