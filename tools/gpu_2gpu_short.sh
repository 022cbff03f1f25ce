#!/bin/bash
# N GPUs, short form: the data-parallel bench line (pipelined step)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_dp$N.json 2> gpurun_out/r2_bench_dp$N.err; echo "bench dp$N rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp$N.json 2>&1 | head -1; grep -v "Warn\|OMP_NUM\|\*\*\*" gpurun_out/r2_bench_dp$N.err | tail -3
