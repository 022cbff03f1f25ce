"""Launch-tuning sweep on one full-size batch (GPU box): times wgrad per layer for every (wg_nbp, wg_lag, wg_occ2)
combination and forward/dgrad with and without the kernel-offset rotation, checking every result against the first
configuration.  Knobs are set through b2s_set_tuning, so one process covers the whole sweep.
Writes gpurun_out/sweep.json."""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import lib, plots
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from dpcr_agb_b200.quantize import GridSampling3D

dev = torch.device("cuda:0")
B = int(os.environ.get("PLOTS", "32"))
b = plots.synth_batch(2, 0, B, n_points=16000)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
gs = GridSampling3D(0.0125)
vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=((0, 0, 0), (80, 80, 100)))
x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], dense_index=vox["index"])
cm = x.coordinate_manager
keys = {1: x.coordinate_map_key}
for ts in (2, 4, 8, 16):
    keys[ts] = cm.stride(keys[ts // 2], 2)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


layers = [("stem k7 3->64 ts1", 1, 1, 7, 3, 64), ("L1 64->64 ts2", 2, 2, 3, 64, 64), ("L2 s2 64->128", 2, 4, 3, 64, 128), ("L2 128->128 ts4", 4, 4, 3, 128, 128),
          ("L3 s2 128->256", 4, 8, 3, 128, 256), ("L3 256->256 ts8", 8, 8, 3, 256, 256),
          ("L4 s2 256->512", 8, 16, 3, 256, 512), ("L4 512->512 ts16", 16, 16, 3, 512, 512)]
# calls per training step of each shape (MSENet14): weights the per-layer times into one figure
calls = {"stem k7 3->64 ts1": 1, "L1 64->64 ts2": 4, "L2 s2 64->128": 1, "L2 128->128 ts4": 3, "L3 s2 128->256": 1, "L3 256->256 ts8": 3,
         "L4 s2 256->512": 1, "L4 512->512 ts16": 3}
prep = []
if os.environ.get("SKIP_STEM"):
    layers = layers[1:]
for name, its, ots, K, cin, cout in layers:
    km = cm.kernel_map(keys[its], keys[ots], K)
    xf = torch.randn(km.n_in, cin, device=dev)
    if cin > 4:
        xf = Fn.round_tf32(xf)
    w = torch.randn(km.k3, cin, cout, device=dev) * 0.02
    gy = Fn.round_tf32(torch.randn(km.n_out, cout, device=dev))
    prep.append((name, km, xf, w, gy, cin, cout))


ALL_KEYS = ("wg_nbp", "wg_lag", "wg_occ2", "wg_ca", "wg_wv", "tc_m256", "tc_rot", "tc_ca", "tc_occ1")


def apply(cfg):
    for k in ALL_KEYS:
        lib.set_tuning(k, cfg.get(k, -1))


res = {"wgrad": [], "fwd": []}
ref = {}
wg_cfgs = [dict(wg_ca=ca, wg_occ2=o2, wg_nbp=nbp) for ca in (0, 1) for o2 in (1, 0) for nbp in (16, 32)]
if os.environ.get("SWEEP_WG"):
    wg_cfgs = json.loads(os.environ["SWEEP_WG"])
for cfg in wg_cfgs:
    apply(cfg)
    row = {"cfg": cfg, "ms": {}, "err": {}}
    tot = 0.0
    for name, km, xf, w, gy, cin, cout in prep:
        f = lambda: Fn.wgrad(xf, gy, km.nbr, km.n_in, km.n_out, cin, cout, km.k3, prerounded=cin > 4)
        g = f()
        if name not in ref:
            ref[name] = g
        row["err"][name] = float((g - ref[name]).abs().max() / ref[name].abs().max())
        t = timeit(f)
        row["ms"][name] = round(t, 4)
        tot += calls[name] * t
    row["step_ms"] = round(tot, 4)
    row["max_err"] = max(row["err"].values())
    del row["err"]
    res["wgrad"].append(row)
    print("wgrad", row, flush=True)

ref = {}
tc_cfgs = [dict(tc_ca=ca, tc_occ1=o1) for ca in (0, 1) for o1 in (0, 1)]
if os.environ.get("SWEEP_TC"):
    tc_cfgs = json.loads(os.environ["SWEEP_TC"])
for cfg in tc_cfgs:
    apply(cfg)
    row = {"cfg": cfg, "fwd_ms": {}, "dgrad_ms": {}}
    tot = 0.0
    err = 0.0
    for name, km, xf, w, gy, cin, cout in prep:
        f = lambda: Fn.gather_gemm(xf, w, None, km.nbr, km.n_in, km.n_out, cin, cout, km.k3, 0, prerounded=cin > 4)
        y = f()
        if name not in ref:
            ref[name] = y
        err = max(err, float((y - ref[name]).abs().max() / ref[name].abs().max()))
        t = timeit(f)
        row["fwd_ms"][name] = round(t, 4)
        tot += calls[name] * t
        if km.symmetric and cin > 4:
            f2 = lambda: Fn.gather_gemm(gy, w, None, km.nbr, km.n_out, km.n_in, cout, cin, km.k3, 3, prerounded=True)
            t2 = timeit(f2)
            row["dgrad_ms"][name] = round(t2, 4)
            tot += calls[name] * t2
    row["step_ms"] = round(tot, 4)
    row["max_err"] = err
    res["fwd"].append(row)
    print("fwd", row, flush=True)
apply({})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/sweep.json", "w"), indent=1)
