#!/bin/bash
# Full state check on the GPU box: gpu test suite, smoke, per-layer conv bench, host profile, bench, ncu launch list.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.jsonl
export B2S_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== full gpu suite ===" | tee gpurun_out/pytest_all.log
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 | tee -a gpurun_out/pytest_all.log
echo "=== smoke ==="
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== conv bench ==="; timeout -k 10 600 python tools/conv_bench.py 2>&1 | tail -20 | tee gpurun_out/conv_bench.log
echo "=== host profile ==="
timeout -k 10 600 python tools/host_profile.py 2>&1 | tail -120 > gpurun_out/host_profile.log; head -3 gpurun_out/host_profile.log
echo "=== bench ==="
timeout -k 10 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_state.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_state.log").read())
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "share_of_step")})
print("cpu", d["cpu_baseline"])
for k, v in d["roofline"]["per_kind"].items(): print(k, v)
tot = 0
for k, v in sorted(d["breakdown_ms_per_step"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]; print(f"  {k:32s} {v['calls_per_step']:5.1f} calls {v['ms_per_step']:8.3f} ms")
print("sum of C-ABI kernels per step", tot)
PY
echo "=== ncu launch list ==="
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
wc -l gpurun_out/launches.csv
