#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
rm -f gpurun_out/parity_report.jsonl
export B2S_PARITY_REPORT=$PWD/gpurun_out/parity_report.jsonl
echo "=== full gpu suite ===" | tee gpurun_out/pytest_all.log
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 | tee -a gpurun_out/pytest_all.log
cat gpurun_out/parity_report.jsonl
echo "=== smoke ==="
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "=== probe ==="; timeout 60 ./tools/umma_probe | tail -2 | tee gpurun_out/probe.log
echo "=== conv bench ==="; timeout -k 10 600 python tools/conv_bench.py 2>&1 | tail -14 | tee gpurun_out/conv_bench.log
echo "=== host profile ==="
timeout -k 10 600 python tools/host_profile.py 2>&1 | tail -120 > gpurun_out/host_profile.log; head -3 gpurun_out/host_profile.log
echo "=== bench ==="
timeout -k 10 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_3.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_3.log").read())
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "share_of_step")})
for k, v in d["roofline"]["per_kind"].items(): print(k, v)
tot = 0
for k, v in sorted(d["breakdown_ms_per_step"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]; print(f"  {k:32s} {v['calls_per_step']:5.1f} calls {v['ms_per_step']:8.3f} ms")
print("sum of C-ABI kernels per step", tot)
PY
