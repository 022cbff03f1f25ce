#!/bin/bash
# 2-GPU bench with and without the overlapped gradient exchange
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for OV in 1 0; do
  B2S_OVERLAP_COMM=$OV timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$OV \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp2_ov$OV.log 2>&1
  echo "overlap=$OV rc=$?"
  grep '^{"metric' gpurun_out/bench_dp2_ov$OV.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','last_loss')}, d['e2e']['value'])"
done
tail -3 gpurun_out/bench_dp2_ov1.log | cut -c1-300
