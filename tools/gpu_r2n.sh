#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2n_bench.json 2>&1 | head -3; tail -3 gpurun_out/r2n_bench.err
timeout 900 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_cfg3.json 2> gpurun_out/r2n_bench_cfg3.err; echo "bench cfg3 rc=$?"
python tools/show_bench.py gpurun_out/r2n_bench_cfg3.json 2>&1 | head -3
B2S_FUSE_BN_STATS=1 timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_graph.py -q -x --timeout 300 2>&1 | tail -3
