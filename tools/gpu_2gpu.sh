#!/bin/bash
# round 2, two GPUs: two-rank == joint-batch gradient test, NCCL bus bandwidth, data-parallel bench line
mkdir -p gpurun_out
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_ddp.py -q --timeout 300 > gpurun_out/r2_ddp_test.log 2>&1; echo "ddp test rc=$?"; tail -5 gpurun_out/r2_ddp_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peak_probe.py 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_dp2.json 2> gpurun_out/r2_bench_dp2.err; echo "bench dp2 rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp2.json 2>&1 | head -3; tail -3 gpurun_out/r2_bench_dp2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_dp2.json 2> gpurun_out/r2_bench_cfg5_dp2.err; echo "bench cfg5 dp2 rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_cfg5_dp2.json 2>&1 | head -3; tail -3 gpurun_out/r2_bench_cfg5_dp2.err
