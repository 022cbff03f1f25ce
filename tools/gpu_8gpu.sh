#!/bin/bash
# one data-parallel bench line at all GPUs of the box (what the driver's scaling run does at round end)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_dp$N.json 2> gpurun_out/r2_bench_dp$N.err; echo "bench dp$N rc=$?"
python tools/show_bench.py gpurun_out/r2_bench_dp$N.json 2>&1 | head -3; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2_bench_dp$N.err | tail -5
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_dp$N.json").read().strip().splitlines()[-1])
print("imbalance", d.get("rank_work_imbalance_max_over_mean"), "ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"])
PY
