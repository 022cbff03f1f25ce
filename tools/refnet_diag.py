"""Diagnostic: gradient of the UNCHANGED reference SENet14 through the product module under different kernel paths
(x-line stem / table-driven tcgen05 / SIMT fp32) against the CPU oracle in fp32 and in fp64."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import test_gpu_reference_nets as T
import dpcr_agb_b200
from dpcr_agb_b200 import lib as L
from dpcr_agb_b200 import plots
from dpcr_agb_b200.MinkowskiEngine import coordinate_manager as CM
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from dpcr_agb_b200.quantize import GridSampling3D
from oracle import me_cpu

cuda = torch.device("cuda:0")
training = os.environ.get("TRAIN", "0") == "1"
b = plots.synth_batch(41, 0, 3, n_points=6000)
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) for k, v in b.items()}
vox = GridSampling3D(0.02)(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=3)
coords, feats = vox["coords"], vox["tensors"][0]


def oracle(dtype):
    me_o, ref = T._reference_senet(me_cpu.install, "SENet14", 0.3)
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    ref = ref.to(dtype)
    ref.train(training)
    random.seed(7)
    yo = ref(me_o.SparseTensor(feats.cpu().to(dtype), coordinates=coords.cpu())).F
    go = torch.from_numpy(np.random.default_rng(5).standard_normal(tuple(yo.shape)).astype(np.float32)).to(dtype)
    yo.backward(go)
    return state, go, {k: p.grad.double().clone() for k, p in ref.named_parameters() if p.grad is not None}, yo.detach().double()


def product(state, go, lines, impl):
    CM.USE_LINES = lines
    Fn.CONV_IMPL = impl
    me_p, net = T._reference_senet(dpcr_agb_b200.install, "SENet14", 0.3)
    net.load_state_dict(state)
    net = net.to(cuda)
    net.train(training)
    random.seed(7)
    yp = net(me_p.SparseTensor(features=feats, coordinates=coords, dense_index=None)).F
    yp.backward(go.float().to(cuda))
    CM.USE_LINES, Fn.CONV_IMPL = True, 0
    return {k: p.grad.detach().cpu().double() for k, p in net.named_parameters() if p.grad is not None}, yp.detach().cpu().double()


state, go, g32, y32 = oracle(torch.float32)
try:
    _, _, g64, y64 = oracle(torch.float64)
except Exception as e:
    print("fp64 oracle failed:", type(e).__name__, e)
    g64, y64 = None, None
runs = {"tc+lines": product(state, go, True, 0), "tc table": product(state, go, False, 0), "simt fp32": product(state, go, False, 1)}


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


names = ["blocks.0.0.conv.kernel", "blocks.0.0.conv.bias", "blocks.1.0.conv1.kernel", "blocks.4.0.conv2.kernel", "final.linear.weight"]
names = [n for n in names if n in g32]
print("training", training, "params checked:", names)
if g64 is not None:
    print("oracle fp32 vs fp64: out", rel(y32, y64), {n: f"{rel(g32[n], g64[n]):.2e}" for n in names})
for tag, (g, y) in runs.items():
    print(tag, "vs oracle fp32: out", f"{rel(y, y32):.2e}", {n: f"{rel(g[n], g32[n]):.2e}" for n in names})
    if g64 is not None:
        print(tag, "vs oracle fp64: out", f"{rel(y, y64):.2e}", {n: f"{rel(g[n], g64[n]):.2e}" for n in names})

# ---- stem wgrad in isolation on the gy the network really produces
print("--- stem wgrad in isolation")
me_p, net = T._reference_senet(dpcr_agb_b200.install, "SENet14", 0.3)
net.load_state_dict(state)
net = net.to(cuda)
net.train(training)
cap = {}
conv = net.conv1[0] if hasattr(net, "conv1") else None
stem = [m for m in net.modules() if isinstance(m, me_p.MinkowskiConvolution)][0]
def _fh(m, i, o):
    o.F.register_hook(lambda g: cap.update(gy=g, tw=getattr(g, "_b2s_tf32", None)))


h = stem.register_forward_hook(_fh)
random.seed(7)
st = me_p.SparseTensor(features=feats, coordinates=coords)
yp = net(st).F
yp.backward(go.float().to(cuda))
gy = cap["gy"]
print("gy", tuple(gy.shape), "contiguous", gy.is_contiguous(), "twin attached", cap["tw"] is not None)
km = st.coordinate_manager.kernel_map(st.coordinate_map_key, st.coordinate_map_key, 7)
nbr = km.nbr
xd, gd = feats.double(), gy.double()
ref = torch.zeros(343, 3, 64, dtype=torch.float64, device=cuda)
for k in range(343):
    o = (nbr[k] >= 0).nonzero().squeeze(1)
    if o.numel():
        ref[k] = xd[nbr[k, o].long()].T @ gd[o]
print("autograd stem grad vs isolated fp64:", rel(stem.kernel.grad.double().view(343, 3, 64), ref))
g_round = Fn.round_tf32(gy)
print("wgrad table (fresh round_tf32):", rel(Fn.wgrad(feats, g_round, nbr, km.n_in, km.n_out, 3, 64, 343, impl=2, prerounded=True).double(), ref))
if cap["tw"] is not None:
    tw = cap["tw"][0]
    print("twin == fresh operand form:", torch.equal(tw, g_round), "max abs diff", (tw - g_round).abs().max().item())
    print("wgrad table (twin):", rel(Fn.wgrad(feats, tw, nbr, km.n_in, km.n_out, 3, 64, 343, impl=2, prerounded=True).double(), ref))
print("wgrad simt:", rel(Fn.wgrad(feats, gy, nbr, km.n_in, km.n_out, 3, 64, 343, impl=1).double(), ref))
print("|gy| max", gy.abs().max().item(), "min nonzero", gy[gy != 0].abs().min().item(), "x max", feats.abs().max(0).values.tolist())
