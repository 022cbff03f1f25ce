#!/bin/bash
# eight GPUs, short form: the data-parallel bench line at 8 and 4 ranks (pipelined step)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_dp8.json 2> gpurun_out/r2_bench_dp8.err; echo "bench dp8 rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp8.json 2>&1 | head -1; grep -v Warn gpurun_out/r2_bench_dp8.err | grep -v "^\*\*\*\|OMP_NUM" | tail -3
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_dp4.json 2> gpurun_out/r2_bench_dp4.err; echo "bench dp4 rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp4.json 2>&1 | head -1
