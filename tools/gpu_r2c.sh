#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/perop_diag.py > gpurun_out/r2c_perop.log 2>&1; echo "perop rc=$?"; cat gpurun_out/r2c_perop.log
PRECISE=1 timeout 600 python tools/conv_bench.py > gpurun_out/r2c_convbench_bf16x2.log 2>&1; echo "rc=$?"
PRECISE=0 timeout 600 python tools/conv_bench.py > gpurun_out/r2c_convbench_tf32.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
a=json.load(open("gpurun_out/conv_bench_bf16x2.json")); b=json.load(open("gpurun_out/conv_bench_tf32.json"))
for la, lb in zip(a["layers"], b["layers"]):
    print(f"{la['layer']:28s} pairs {la['pairs']:9d} map {la['kernel_map_ms']:.3f} | " + "  ".join(
        f"{k[:-3]} {lb.get(k,0):.3f}->{la.get(k,0):.3f}" for k in ("fwd_ms","dgrad_ms","dgrad_parity_ms","wgrad_ms") if k in la))
PY
