#!/bin/bash
# two GPUs, short form: two-rank == joint-batch gradient test + the data-parallel bench line (pipelined step)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddp.py -q --timeout 250 > gpurun_out/r2_ddp_test.log 2>&1; echo "ddp test rc=$?"; tail -3 gpurun_out/r2_ddp_test.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_dp2.json 2> gpurun_out/r2_bench_dp2.err; echo "bench dp2 rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp2.json 2>&1 | head -3; grep -v Warn gpurun_out/r2_bench_dp2.err | tail -3
B2S_PIPELINE=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_dp2_nopipe.json 2> gpurun_out/r2_bench_dp2_nopipe.err; echo "bench dp2 (single graph) rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp2_nopipe.json 2>&1 | head -1
