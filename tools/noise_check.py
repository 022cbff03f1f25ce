"""Run-to-run spread of the training loss of ONE model on ONE small batch (2 plots x 1200 points: split-K launches
with fp32 reductions in arrival order, batch norm over a handful of rows at the coarse levels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import msenet, train
import test_gpu_f4 as t4

dev = torch.device("cuda:0")
batch, c, f = t4._voxels(num_plots=2, n_points=1200)
cg, fg = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)
target = torch.from_numpy(batch["target"]).to(dev)
torch.manual_seed(0)
tr = train.Trainer(msenet.build(ME, "SENet14", drop_path=0.0).to(dev), ME, lr=1e-3)
for _ in range(2):
    tr.step(cg, fg, target)
vals = []
with torch.no_grad():
    for _ in range(8):
        x = ME.SparseTensor(features=fg, coordinates=cg)
        vals.append(float(train.reg_loss(tr.model(x), target, tr.center, tr.scale)))
v = np.array(vals)
print("losses", vals)
print("relative spread (max - min) / mean = %.3e" % ((v.max() - v.min()) / v.mean()))
