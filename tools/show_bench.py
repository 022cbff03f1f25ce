import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "frac", "share_of_step", "bound", "unit")})
print("clocks", d.get("clocks"))
print("cpu", d.get("cpu_baseline"))
for k, v in d["roofline"].get("per_kind", {}).items(): print(k, v)
tot = 0
for k, v in sorted(d.get("breakdown_ms_per_step", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]; print(f"  {k:32s} {v['calls_per_step']:5.1f} calls {v['ms_per_step']:8.3f} ms")
print("sum of C-ABI kernels per step", tot)
