"""Small driver for `ncu --set full`: builds the maps of one full-size batch and runs the hot kernels a few times
(stem map + stem conv, L1 64->64 forward / dgrad / wgrad, L2 128->128 forward / wgrad).  Not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import plots
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from dpcr_agb_b200.quantize import GridSampling3D

dev = torch.device("cuda:0")
B = int(os.environ.get("PLOTS", "32"))
b = plots.synth_batch(2, 0, B, n_points=16000)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
gs = GridSampling3D(0.0125)
reps = int(os.environ.get("REPS", "3"))
for _ in range(reps):
    vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=((0, 0, 0), (80, 80, 100)))
    x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"])
    cm = x.coordinate_manager
    k1 = x.coordinate_map_key
    k2 = cm.stride(k1, 2)
    k4 = cm.stride(k2, 2)
    km7 = cm.kernel_map(k1, k1, 7)
    w7 = torch.randn(343, 3, 64, device=dev) * 0.02
    y = Fn.gather_gemm(vox["tensors"][0], w7, None, km7.nbr, km7.n_in, km7.n_out, 3, 64, 343, 0)
    Fn.wgrad(vox["tensors"][0], Fn.round_tf32(torch.randn_like(y)), km7.nbr, km7.n_in, km7.n_out, 3, 64, 343, prerounded=True)
    for key, c in ((k2, 64), (k4, 128)):
        km = cm.kernel_map(key, key, 3)
        xf = Fn.round_tf32(torch.randn(km.n_in, c, device=dev))
        gy = Fn.round_tf32(torch.randn(km.n_out, c, device=dev))
        w = torch.randn(27, c, c, device=dev) * 0.02
        Fn.gather_gemm(xf, w, None, km.nbr, km.n_in, km.n_out, c, c, 27, 0, prerounded=True)
        Fn.gather_gemm(gy, w, None, km.nbr, km.n_out, km.n_in, c, c, 27, 3, prerounded=True)
        Fn.wgrad(xf, gy, km.nbr, km.n_in, km.n_out, c, c, 27, prerounded=True)
    km_s = cm.kernel_map(k2, k4, 3)
    gy = Fn.round_tf32(torch.randn(km_s.n_out, 128, device=dev))
    w = torch.randn(27, 64, 128, device=dev) * 0.02
    Fn.gather_gemm(gy, w, None, km_s.inv, km_s.n_out, km_s.n_in, 128, 64, 27, 1, prerounded=True)
torch.cuda.synchronize()
print("done")
