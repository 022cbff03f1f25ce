#!/bin/bash
# session 9, call B: 2-GPU data-parallel bench under torchrun (NCCL all-reduce inside the captured step graph)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/bench_dp2.log | cut -c1-1500
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_dp2.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/bench_ref_dp2.log | cut -c1-800
