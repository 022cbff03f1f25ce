"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) of bench.py's timed graph
replays: one training step = the launches from one adabelief_kernel launch to the next; per kernel (template arguments
kept, parameter list dropped) the launches per step, summed duration and share of the step.  ncu serialises the
kernels and runs them cold-cache: compare SHARES with bench.py's CUDA-event breakdown, not absolute times.
usage: python tools/ncu_launch_summary.py launches.csv [steps]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, steps=2):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rd:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]]
        unit = r[idx["Metric Unit"]]
        val = float(r[idx["Metric Value"]].replace(",", ""))
        val *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
        short = re.sub(r"\(.*$", "", name).replace("void ", "").replace("<unnamed>::", "")
        short = re.sub(r"^.*?::(?=[a-z_0-9]+(<|$))", "", short) if short.startswith("at::") else short
        rows.append((short, val))
    opt = [i for i, (n, _) in enumerate(rows) if "adabelief_kernel" in n]
    print(f"{len(rows)} launches captured, {len(opt)} optimiser launches (= steps)")
    if len(opt) >= 2:
        step = rows[opt[-2] + 1: opt[-1] + 1]
    else:
        step = rows
    total = sum(v for _, v in step)
    print(f"one training step = the launches between two adabelief_kernel launches: {len(step)} launches, "
          f"{total:.3f} ms summed (cold-cache, serialised: compare SHARES)")
    agg = OrderedDict()
    for n, v in step:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + v)
    conv = sum(t for n, (c, t) in agg.items() if re.search(r"gather_gemm|wgrad|conv_lines", n))
    print(f"convolution kernels (gather_gemm_tc*, wgrad_*, conv_lines_*): {conv:.3f} ms = {100 * conv / total:.1f}% of the step")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:76]:76s} {c:4d} {t:9.3f} ms {100 * t / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
