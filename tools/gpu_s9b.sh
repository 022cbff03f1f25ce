#!/bin/bash
# 2-GPU data-parallel bench under torchrun (NCCL all-reduce inside the captured step graph) + reference arm at N=2
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
( time timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2.log 2>&1 ) 2>&1 | grep real
echo "rc=$?"; grep '^{"metric' gpurun_out/bench_dp2.log | cut -c1-400; tail -3 gpurun_out/bench_dp2.log | cut -c1-300
