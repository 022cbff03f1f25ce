"""Timing of the k3 s2 max pooling (forward + backward) behind the stem on one full-size batch (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import lib as L
from dpcr_agb_b200 import plots
from dpcr_agb_b200.quantize import GridSampling3D

dev = torch.device("cuda:0")
B = 32
b = plots.synth_batch(2, 0, B, n_points=16000)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
vox = GridSampling3D(0.0125)(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B,
                             bounds=((0, 0, 0), (80, 80, 100)))
x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], dense_index=vox["index"])
cm = x.coordinate_manager
k2 = cm.stride(x.coordinate_map_key, 2)
km = cm.kernel_map(x.coordinate_map_key, k2, 3)
c = 64
f = torch.randn(km.n_in, c, device=dev)
y = torch.empty((km.n_out, c), device=dev)
arg = torch.empty((km.n_out, c), dtype=torch.int32, device=dev)
yr = torch.empty_like(y)


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


pairs = km.num_pairs()
t = timeit(lambda: L.call("b2s_maxpool_fwd", f, km.nbr, km.n_out, km.n_out_dev, c, km.k3, y, arg, yr))
byts = 4 * c * (pairs + 3 * km.n_out) + 4 * km.k3 * km.n_out
print(f"maxpool fwd: rows {km.n_in} -> {km.n_out}, pairs {pairs}: {t:.1f} us, {byts / t / 1e3:.0f} GB/s of touched bytes")
