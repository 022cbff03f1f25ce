"""Diagnostic (GPU): per-tensor end-to-end gradient error of the SIMT fp32 path and of the tcgen05 path (both operand
modes) against the fp32 oracle at BASELINE plot size -- separates the fp32 noise floor of the problem (summation
order amplified by training-mode batch norm) from operand precision.  Writes gpurun_out/parity_diag.json."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import b2s_testutil as util
from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import lib, msenet, train
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from oracle import me_cpu
from oracle import train as otrain

dev = torch.device("cuda:0")
n_plots = int(sys.argv[1]) if len(sys.argv) > 1 else 4
batch = util.make_points(n_plots, 16000, cfg=2)
c, f, _, _, _ = util.oracle_quantize(batch, 0.0125)
target = torch.from_numpy(batch["target"])
center, scale = torch.tensor([107.0, 200.0]), torch.tensor([103.0, 194.0])
torch.manual_seed(0)
ref = msenet.MSENet(me_cpu, "SENet14", drop_path=0.2)
ref.train()
random.seed(11)
yr = ref(me_cpu.SparseTensor(torch.from_numpy(f), coordinates=torch.from_numpy(c)))
otrain.reg_loss(yr, target, center, scale).backward()
gref = {n: p.grad.double() for n, p in ref.named_parameters() if p.grad is not None}
gmax = max(g.abs().max().item() for g in gref.values())
out = {}
for tag, impl, precise in (("simt", 1, 1), ("tc_bf16x2", 0, 1), ("tc_tf32", 0, 0)):
    Fn.CONV_IMPL = impl
    lib.set_tuning("precise", precise)
    mine = msenet.MSENet(ME, "SENet14", drop_path=0.2)
    mine.load_state_dict({k: v.clone() for k, v in ref.state_dict().items()})
    # the oracle's BN buffers have been updated by its forward: restore the initial ones
    for b in mine.buffers():
        if b.dtype.is_floating_point:
            b.copy_(torch.ones_like(b) if b.mean() > 0.5 else torch.zeros_like(b))
    mine = mine.to(dev).train()
    random.seed(11)
    ym = mine(ME.SparseTensor(features=torch.from_numpy(f), coordinates=torch.from_numpy(c), device=dev))
    train.reg_loss(ym, target.to(dev), center.to(dev), scale.to(dev)).backward()
    rows = {}
    for n, p in mine.named_parameters():
        if n not in gref or gref[n].abs().max().item() < 1e-5 * gmax:
            continue
        a, b = p.grad.double().cpu(), gref[n]
        bm = b.abs().max().item()
        d = (a - b).abs()
        rows[n] = {"inf": (d.max() / bm).item(),
                   "el_0.1": (d / torch.clamp(b.abs(), min=0.1 * bm)).max().item(),
                   "el_0.01": (d / torch.clamp(b.abs(), min=0.01 * bm)).max().item()}
    worst = sorted(rows.items(), key=lambda kv: -kv[1]["inf"])[:4]
    out[tag] = {"out": util.rel_err(ym, yr), "worst": worst,
                "worst_el_0.1": max(r["el_0.1"] for r in rows.values()),
                "worst_el_0.01": max(r["el_0.01"] for r in rows.values()),
                "stem_kernel": rows.get("blocks.0.0.conv.kernel")}
    print(tag, json.dumps(out[tag]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_diag.json"), "w"), indent=1)
