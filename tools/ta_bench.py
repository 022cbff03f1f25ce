"""A/B of the convolution forward / dgrad kernels with the A operand staged in shared memory (tc_ta = 0) and gathered
into tensor memory (tc_ta = 1, 2) on one full-size batch (GPU box).  Prints per layer ms and algorithmic TFLOP/s."""
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
for ts in (2, 4, 8):
    keys[ts] = cm.stride(keys[ts // 2], 2)


def timeit(fn, reps=int(os.environ.get('TA_REPS', '10'))):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, ts, cin, cout in (("L1 64->64 @ts2", 2, 64, 64), ("L2 128->128 @ts4", 4, 128, 128),
                            ("64->128 @ts2", 2, 64, 128), ("128->64 @ts2", 2, 128, 64))[:int(os.environ.get("TA_LAYERS", "4"))]:
    km = cm.kernel_map(keys[ts], keys[ts], 3)
    pairs = km.num_pairs()
    flops = 2.0 * pairs * cin * cout
    xf = Fn.round_tf32(torch.randn(km.n_in, cin, device=dev))
    gy = Fn.round_tf32(torch.randn(km.n_out, cout, device=dev))
    w = torch.randn(km.k3, cin, cout, device=dev) * 0.02
    ref = None
    for mode in [int(m) for m in os.environ.get("TA_MODES", "0,1,2").split(",")]:
        lib.set_tuning("tc_ta", mode)
        fwd = lambda: Fn.gather_gemm(xf, w, None, km.nbr, km.n_in, km.n_out, cin, cout, km.k3, 0, prerounded=True)
        dg = lambda: Fn.gather_gemm(gy, w, None, km.nbr, km.n_out, km.n_in, cout, cin, km.k3, 3, prerounded=True)
        y, gx = fwd(), dg()
        if ref is None:
            ref = (y, gx)
        err = max(((y - ref[0]).abs().max() / ref[0].abs().max()).item(),
                  ((gx - ref[1]).abs().max() / ref[1].abs().max()).item())
        tf, td = timeit(fwd), timeit(dg)
        print(f"{name:18s} rows {km.n_out:7d} tc_ta={mode}: fwd {tf * 1e3:7.1f} us {flops / tf / 1e9:6.1f} TF/s | "
              f"dgrad {td * 1e3:7.1f} us {flops / td / 1e9:6.1f} TF/s | max rel diff vs tc_ta=0 {err:.2e}", flush=True)
    lib.set_tuning("tc_ta", -1)
