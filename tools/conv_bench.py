"""Per-layer timing of the MSENet14 convolution stack on one full-size batch (GPU box): kernel maps, forward,
dgrad, wgrad of every distinct conv shape, with algorithmic TFLOP/s.  Writes gpurun_out/conv_bench.json."""
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
PRECISE = int(os.environ.get("PRECISE", "1"))      # operand mode: 1 = split-bf16 (default), 0 = TF32
lib.set_tuning("precise", PRECISE)
B = int(os.environ.get("PLOTS", "32"))
b = plots.synth_batch(2, 0, B, n_points=16000)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
gs = GridSampling3D(0.0125)


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


res = {}
res["quantize_ms"] = timeit(lambda: gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B,
                                       bounds=((0, 0, 0), (80, 80, 100))))
vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=((0, 0, 0), (80, 80, 100)))
x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"],
                    dense_index=None if os.environ.get("NO_DENSE") else vox["index"])
cm = x.coordinate_manager
keys = {1: x.coordinate_map_key}
for ts in (2, 4, 8, 16):
    keys[ts] = cm.stride(keys[ts // 2], 2)
rows = {ts: cm.coords(k).shape[0] for ts, k in keys.items()}
res["rows"] = rows
print("rows per level", rows)


def build_map(in_ts, out_ts, K):
    cm.kernel_maps.clear()
    return cm.kernel_map(keys[in_ts], keys[out_ts], K)


layers = [("stem k7 3->64 @ts1", 1, 1, 7, 3, 64), ("pool k3s2 ts1->2", 1, 2, 3, 0, 0),
          ("L1 k3 64->64 @ts2", 2, 2, 3, 64, 64), ("L2 k3s2 64->128 ts2->4", 2, 4, 3, 64, 128),
          ("L2 k1s2 64->128 ts2->4", 2, 4, 1, 64, 128), ("L2 k3 128->128 @ts4", 4, 4, 3, 128, 128),
          ("L3 k3s2 128->256 ts4->8", 4, 8, 3, 128, 256), ("L3 k3 256->256 @ts8", 8, 8, 3, 256, 256),
          ("L4 k3s2 256->512 ts8->16", 8, 16, 3, 256, 512), ("L4 k3 512->512 @ts16", 16, 16, 3, 512, 512)]
out = []
if os.environ.get("ONLY_STEM"):
    layers = layers[:1]
for name, its, ots, K, cin, cout in layers:
    t_map = timeit(lambda: build_map(its, ots, K), reps=3)
    km = build_map(its, ots, K)
    lines = cin and Fn.lines_path(km, cin, cout)
    if lines:                                   # x-line form: what the stem really builds per step
        def build_lines():
            cm.kernel_maps.clear()
            return cm.kernel_map(keys[its], keys[ots], K).lines
        t_map = timeit(build_lines, reps=3)
        _ = km.lines
    pairs = km.num_pairs()
    row = {"layer": name, "n_in": km.n_in, "n_out": km.n_out, "k3": km.k3, "pairs": pairs,
           "fill": pairs / (km.n_out * km.k3), "kernel_map_ms": t_map,
           "kernel_map_GBps": (16 * km.n_out + 4 * km.k3 * km.n_out) / t_map / 1e6}
    if lines:
        row["kernel_map_GBps"] = (16 * km.n_out + 4 * km.kernel_size[1] * km.kernel_size[2] * km.n_out) / t_map / 1e6
    if cin:
        xf = Fn.round_tf32(torch.randn(km.n_in, cin, device=dev))          # operands pre-rounded, as the autograd
        w = torch.randn(km.k3, cin, cout, device=dev) * 0.02                # Function hands them to the kernels
        gy = Fn.round_tf32(torch.randn(km.n_out, cout, device=dev))
        pre = cin > 4
        flops = 2.0 * pairs * cin * cout
        if lines:
            xp = torch.randn(km.n_in, cin, device=dev)
            t = timeit(lambda: Fn.lines_fwd(xp, w, None, km, cin, cout))
            row.update(fwd_ms=t, fwd_tflops=flops / t / 1e9, path="x-lines")
            t = timeit(lambda: Fn.lines_wgrad(xp, gy, km, cin, cout))
            row.update(wgrad_ms=t, wgrad_tflops=flops / t / 1e9)
            out.append(row)
            print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()})
            continue
        t = timeit(lambda: Fn.gather_gemm(xf, w, None, km.nbr, km.n_in, km.n_out, cin, cout, km.k3, 0, prerounded=pre))
        row.update(fwd_ms=t, fwd_tflops=flops / t / 1e9)
        if cin >= 32:
            tbl = km.nbr if km.symmetric else km.inv
            lay = 3 if km.symmetric else 1
            t = timeit(lambda: Fn.gather_gemm(gy, w, None, tbl, km.n_out, km.n_in, cout, cin, km.k3, lay, prerounded=True))
            row.update(dgrad_ms=t, dgrad_tflops=flops / t / 1e9)
            if not km.symmetric and km.parity_plan is not None:
                t = timeit(lambda: Fn.dgrad_strided(gy, w, km, cout, cin))
                row.update(dgrad_parity_ms=t, dgrad_parity_tflops=flops / t / 1e9)
        t = timeit(lambda: Fn.wgrad(xf, gy, km.nbr, km.n_in, km.n_out, cin, cout, km.k3, prerounded=True))
        row.update(wgrad_ms=t, wgrad_tflops=flops / t / 1e9)
    out.append(row)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()})
res["layers"] = out
os.makedirs("gpurun_out", exist_ok=True)
res["precise"] = PRECISE
json.dump(res, open(f"gpurun_out/conv_bench_{'bf16x2' if PRECISE else 'tf32'}.json", "w"), indent=1)
