"""Driver for `ncu --set full`: runs one instance of every hot-path stage on one full-size batch, each instance
preceded by a MARKER launch (a one-element b2s_gelu_fwd == flat_kernel with grid 1), and writes a manifest with the
ALGORITHMIC bytes / FLOPs of every instance (SURVEY.md 8d formulas).  tools/ncu_instances.py joins the ncu export with
the manifest by splitting the launch list at the markers.  Not a benchmark."""
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
PRECISE = int(os.environ.get("PRECISE", "1"))
lib.set_tuning("precise", PRECISE)
b = plots.synth_batch(2, 0, B, n_points=16000)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in b.items()}
gs = GridSampling3D(0.0125)
BOUNDS = ((0, 0, 0), (80, 80, 100))
manifest = []
one = torch.zeros(1, device=dev)


def instance(tag, fn, bound, alg_bytes=None, alg_flops=None, note=""):
    lib.call("b2s_gelu_fwd", one, 1, None, 1, one)          # marker
    out = fn()
    manifest.append({"tag": tag, "bound": bound, "alg_bytes": alg_bytes, "alg_flops": alg_flops, "note": note})
    return out


def conv_bytes(n_in, n_out, cin, cout, k3, pairs):
    return 4 * (n_in * cin + n_out * cout) + 4 * k3 * cin * cout + 8 * pairs


# warm-up (kernel attributes, allocator) -- not captured: the ncu command skips everything before the first marker
for _ in range(2):
    gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=BOUNDS)
n = d["pos"].shape[0]
vox = instance("quantize (GridSampling3D: points -> voxels, 3-feature gather x2)",
               lambda: gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=B, bounds=BOUNDS),
               "hbm", note="12n + 4n (perm) + 12M + 8M + feature gather 4F(M+M), F = 3 (x) and 3 (pos)")
M = vox["coords"].shape[0]
manifest[-1]["alg_bytes"] = 12 * n + 4 * n + 12 * M + 8 * M + 2 * 4 * 3 * (M + M)
x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], dense_index=vox["index"])
cm = x.coordinate_manager
k1 = x.coordinate_map_key
k2 = instance("strided map ts1 -> ts2 (hash insert + fill)", lambda: cm.stride(k1, 2), "hbm")
N2 = cm.coords(k2).shape[0]
manifest[-1]["alg_bytes"] = 16 * M + 16 * N2 + 4 * M
k4 = cm.stride(k2, 2)
LINES = Fn.lines_path(cm.kernel_map(k1, k1, 7), 3, 64)
cm.kernel_maps.clear()
if LINES:
    km7 = instance("kernel map k7 stem (occupancy index, x-line table)", lambda: cm.kernel_map(k1, k1, 7), "hbm")
    manifest[-1]["kernels_expected"] = 0            # the map object is lazy: the table is the next instance's launch
    instance("x-line table of the k7 stem map", lambda: km7.lines, "hbm")
    P7 = km7.num_pairs()
    manifest[-1]["alg_bytes"] = 16 * M + 8 * P7
    manifest[-1]["note"] = f"16 N_out + 8 P (P = {P7}); the kernel WRITES 4 * 49 * N_out = {4 * 49 * M} bytes of line words"
else:
    km7 = instance("kernel map k7 stem (occupancy index, dense table)", lambda: cm.kernel_map(k1, k1, 7), "hbm")
    P7 = int((km7.nbr >= 0).sum())
    manifest[-1]["alg_bytes"] = 16 * M + 8 * P7
    manifest[-1]["note"] = f"16 N_out + 8 P (P = {P7}); the kernel WRITES the dense table 4 * 343 * N_out = {4 * 343 * M} bytes"
km2 = instance("kernel map k3 @ts2 (hash probes)", lambda: cm.kernel_map(k2, k2, 3), "hbm")
P2 = int((km2.nbr >= 0).sum())
manifest[-1]["alg_bytes"] = 16 * N2 + 8 * P2
w7 = torch.randn(343, 3, 64, device=dev) * 0.02
f1 = vox["tensors"][0]
if LINES:
    y = instance("conv fwd stem k7 3->64 @ts1 (x-lines)", lambda: Fn.lines_fwd(f1, w7, None, km7, 3, 64), "tensor",
                 conv_bytes(M, M, 3, 64, 343, P7), 2 * P7 * 3 * 64)
    gy1 = Fn.round_tf32(torch.randn_like(y))
    instance("conv wgrad stem k7 3->64 @ts1 (x-lines)", lambda: Fn.lines_wgrad(f1, gy1, km7, 3, 64), "tensor",
             conv_bytes(M, M, 3, 64, 343, P7), 2 * P7 * 3 * 64)
else:
    y = instance("conv fwd stem k7 3->64 @ts1", lambda: Fn.gather_gemm(f1, w7, None, km7.nbr, M, M, 3, 64, 343, 0),
                 "tensor", conv_bytes(M, M, 3, 64, 343, P7), 2 * P7 * 3 * 64)
    gy1 = Fn.round_tf32(torch.randn_like(y))
    instance("conv wgrad stem k7 3->64 @ts1", lambda: Fn.wgrad(f1, gy1, km7.nbr, M, M, 3, 64, 343, prerounded=True),
             "tensor", conv_bytes(M, M, 3, 64, 343, P7), 2 * P7 * 3 * 64)
if os.environ.get("ONLY_STEM"):
    lib.call("b2s_gelu_fwd", one, 1, None, 1, one)
    torch.cuda.synchronize()
    json.dump({"plots": B, "precise": PRECISE, "rows": {"ts1": M, "ts2": N2}, "instances": manifest},
              open("gpurun_out/ncu_manifest.json", "w"), indent=1)
    print("done", len(manifest), "instances")
    sys.exit(0)
for key, c, name in ((k2, 64, "L1 k3 64->64 @ts2"), (k4, 128, "L2 k3 128->128 @ts4")):
    km = cm.kernel_map(key, key, 3)
    P = int((km.nbr >= 0).sum())
    xf = Fn.round_tf32(torch.randn(km.n_in, c, device=dev))
    gy = Fn.round_tf32(torch.randn(km.n_out, c, device=dev))
    w = torch.randn(27, c, c, device=dev) * 0.02
    cb, fl = conv_bytes(km.n_in, km.n_out, c, c, 27, P), 2 * P * c * c
    instance(f"conv fwd {name}", lambda: Fn.gather_gemm(xf, w, None, km.nbr, km.n_in, km.n_out, c, c, 27, 0,
                                                         prerounded=True), "tensor", cb, fl)
    instance(f"conv dgrad {name}", lambda: Fn.gather_gemm(gy, w, None, km.nbr, km.n_out, km.n_in, c, c, 27, 3,
                                                           prerounded=True), "tensor", cb, fl)
    instance(f"conv wgrad {name}", lambda: Fn.wgrad(xf, gy, km.nbr, km.n_in, km.n_out, c, c, 27, prerounded=True),
             "tensor", cb, fl)
km_s = cm.kernel_map(k2, k4, 3)
Ps = int((km_s.nbr >= 0).sum())
_ = km_s.parity_plan
gys = Fn.round_tf32(torch.randn(km_s.n_out, 128, device=dev))
ws = torch.randn(27, 64, 128, device=dev) * 0.02
instance("conv dgrad (parity plan) L2 k3s2 64->128 ts2->4", lambda: Fn.dgrad_strided(gys, ws, km_s, 128, 64), "tensor",
         conv_bytes(km_s.n_in, km_s.n_out, 64, 128, 27, Ps), 2 * Ps * 64 * 128)
xb = torch.randn(N2, 64, device=dev)
mean, invstd = torch.zeros(64, device=dev), torch.ones(64, device=dev)
wsb = torch.empty(2 * 64 + 1, dtype=torch.float64, device=dev)
instance("batch-norm statistics [N2, 64]", lambda: lib.call("b2s_bn_stats", xb, N2, None, 64, 1e-5, 0.1, None, None, wsb,
                                                            mean, invstd), "hbm", 4 * N2 * 64)
yb, ybr = torch.empty_like(xb), torch.empty_like(xb)
instance("batch-norm apply + GELU (+ operand twin) [N2, 64]",
         lambda: lib.call("b2s_bn_apply", xb, mean, invstd, None, None, N2, None, 64, 1, yb, ybr), "hbm", 8 * N2 * 64,
         note="8 N C algorithmic (read + write); the operand twin is a second write")
lib.call("b2s_gelu_fwd", one, 1, None, 1, one)              # closing marker
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"plots": B, "precise": PRECISE, "rows": {"ts1": M, "ts2": N2}, "instances": manifest},
          open("gpurun_out/ncu_manifest.json", "w"), indent=1)
print("done", len(manifest), "instances")
