#!/bin/bash
# round 2, run B: parity diagnostic (fp32 noise floor vs operand precision) + bench in both operand modes
mkdir -p gpurun_out
timeout 600 python tools/parity_diag.py 4 > gpurun_out/r2b_diag.log 2>&1; echo "diag rc=$?"; cat gpurun_out/r2b_diag.log | cut -c1-1500
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision bf16x2 > gpurun_out/r2b_bench_bf16x2.json 2> gpurun_out/r2b_bench_bf16x2.err; echo "bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --precision tf32 > gpurun_out/r2b_bench_tf32.json 2> gpurun_out/r2b_bench_tf32.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2b_bench_bf16x2.json gpurun_out/r2b_bench_tf32.json
