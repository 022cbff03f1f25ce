"""Join an `ncu --page raw --csv` export of tools/ncu_target.py with its manifest: the launch list is split at the
marker launches (flat_kernel with a grid of one CTA); per instance it reports duration, DRAM traffic (dram__bytes_read
+ dram__bytes_write summed over the instance's kernels), tensor-pipe activity of its longest kernel, and the achieved
fraction of the roofline from the manifest's ALGORITHMIC bytes / FLOPs.
usage: python tools/ncu_instances.py raw.csv manifest.json out.json [MEASURED_PEAKS.json]"""
import csv
import json
import sys

UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
        "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main(raw, manifest_path, out_path, peaks_path=None):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        if key not in idx or r[idx[key]] == "":
            return None
        return float(r[idx[key]].replace(",", "")) * UNIT.get(units[idx[key]], 1.0)

    man = json.load(open(manifest_path))
    peaks = json.load(open(peaks_path)) if peaks_path else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    bf16 = peaks.get("bf16_tflops", 1590.0)          # burst figure: every kernel here is timed alone
    tensor_peak = bf16 / (3.0 if man.get("precise", 1) else 2.0)
    groups, cur = [], None
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        grid = r[idx["launch__grid_size"]] if "launch__grid_size" in idx else ""
        if "flat_kernel" in name and grid.replace(",", "").split(".")[0] == "1":
            cur = []
            groups.append(cur)
        elif cur is not None and "at::" not in name:      # torch's own kernels (statistics code of the driver script)
            cur.append(r)
    inst = man["instances"]
    if len(groups) == len(inst) + 1 and not groups[-1]:
        groups = groups[:-1]                                  # the closing marker opens an empty group
    # an instance may launch nothing (a lazily built map): keep its (empty) group, drop it from the report
    if len(groups) == len(inst):
        keep = [i for i, g in enumerate(groups) if g]
        groups, inst = [groups[i] for i in keep], [inst[i] for i in keep]
    else:
        groups = [g for g in groups if g]
        inst = [m for m in inst if m.get("kernels_expected", 1) != 0]
    assert len(groups) == len(inst), f"{len(groups)} marker-delimited groups for {len(inst)} manifest instances"
    out = []
    for m, g in zip(inst, groups):
        t = sum(val(r, "gpu__time_duration.sum") or 0.0 for r in g)
        dram = sum((val(r, "dram__bytes_read.sum") or 0.0) + (val(r, "dram__bytes_write.sum") or 0.0) for r in g)
        top = max(g, key=lambda r: val(r, "gpu__time_duration.sum") or 0.0)
        e = dict(m)
        e.update({"kernels": [r[idx["Kernel Name"]].split("(")[0][:70] for r in g], "time_us": t * 1e6,
                  "dram_bytes": dram,
                  "dram_over_algorithmic": dram / m["alg_bytes"] if m.get("alg_bytes") else None,
                  "top_kernel_tensor_pipe_pct": val(top, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                  "top_kernel_dram_pct": val(top, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                  "top_kernel_regs": val(top, "launch__registers_per_thread")})
        if m["bound"] == "hbm" and m.get("alg_bytes"):
            e["achieved_GBps"] = m["alg_bytes"] / t / 1e9
            e["frac_of_measured_hbm"] = e["achieved_GBps"] / hbm
            e["dram_GBps"] = dram / t / 1e9
        if m["bound"] == "tensor" and m.get("alg_flops"):
            e["achieved_TFLOPs"] = m["alg_flops"] / t / 1e12
            e["frac_of_tensor_roofline"] = e["achieved_TFLOPs"] / tensor_peak
        out.append(e)
    json.dump({"source": "ncu --set full --clock-control none over tools/ncu_target.py (cold-cache, serialised)",
               "precise": man.get("precise", 1), "peaks": {"hbm_GBps": hbm, "tensor_TFLOPs": tensor_peak,
                                                             "tensor_note": "measured bf16 burst / 3 products (split-bf16) "
                                                             "or / 2 (tf32 rate)"},
               "instances": out}, open(out_path, "w"), indent=1)
    for e in out:
        extra = (f"{e['achieved_GBps']:7.0f} GB/s alg ({100 * e['frac_of_measured_hbm']:4.1f}% of HBM), dram {e['dram_GBps']:6.0f} GB/s"
                 if "achieved_GBps" in e else
                 f"{e.get('achieved_TFLOPs', 0):6.1f} TF/s alg ({100 * e.get('frac_of_tensor_roofline', 0):4.1f}%), tensor pipe {e['top_kernel_tensor_pipe_pct'] or 0:4.1f}%")
        ratio = f"{e['dram_over_algorithmic']:.2f}x" if e["dram_over_algorithmic"] else "-"
        print(f"{e['tag'][:58]:58s} {e['time_us']:8.1f} us  dram {e['dram_bytes'] / 1e6:8.1f} MB ({ratio} alg)  {extra}")


if __name__ == "__main__":
    main(*sys.argv[1:])
