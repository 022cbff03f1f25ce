"""Top SASS instructions by warp-stall samples from `ncu --page source --csv`, per kernel in the export, with the
dominant stall reasons per instruction and the totals per reason.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kernels, cur = [], None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif r[0] == "Address" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
for k in kernels:
    hdr = k["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data, tot, total = [], {h: 0 for h in stall_cols}, 0
    for r in k["rows"]:
        n = int(r[col["# Samples"]] or 0)
        total += n
        st = {h: int(r[col[h]] or 0) for h in stall_cols}
        for h, v in st.items():
            tot[h] += v
        data.append((n, r[col["Source"]].strip(), st, int(r[col["Instructions Executed"]] or 0)))
    print("==", k["name"][:120])
    print("total samples", total)
    print("by reason:", ", ".join(f"{h[6:]}={v} ({100 * v / max(total, 1):.0f}%)"
                                  for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
    for n, src, st, ex in sorted(data, key=lambda d: -d[0])[:top_n]:
        top = ", ".join(f"{h[6:]}={v}" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{n:7d} {100 * n / max(total, 1):5.1f}%  exec={ex:9d}  {src[:70]:70s} {top}")
