"""Pretty-print the last JSON line of a bench.py output file."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d.get("metric"), {k: d[k] for k in ("value", "ms_per_step", "gpu_launches") if k in d}, "e2e", d["e2e"]["value"],
      d["e2e"].get("ms_per_step"))
if "roofline" in d:
    r = d["roofline"]
    print("roofline fwd+dgrad", {k: r.get(k) for k in ("achieved", "frac", "share_of_step", "executed_frac_of_bf16_peak", "traffic")})
    w = d.get("roofline_wgrad") or {}
    print("roofline wgrad    ", {k: w.get(k) for k in ("achieved", "frac", "share_of_step", "ms_per_step")})
    for s in (d.get("roofline_hbm") or {}).get("stages", []):
        print(f"  hbm {s['stage'][:48]:48s} {s['ms_per_step']:.3f} ms {s['algorithmic_mb_per_step']:8.1f} MB "
              f"{s['achieved']:7.0f} GB/s  frac {s['frac']:.3f}")
    for k, v in r.get("per_kind", {}).items():
        print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "entry_points"})
print("clocks", d.get("clocks"))
print("cpu", d.get("cpu_baseline"))
tot = 0
for k, v in sorted(d.get("breakdown_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    print(f"  {k:32s} {v['calls_per_step']:5.1f} calls {v['ms_per_step']:8.3f} ms")
print("sum of C-ABI kernels per step", tot)
