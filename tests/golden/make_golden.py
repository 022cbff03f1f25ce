"""Regenerates the committed fixtures under tests/golden/.  Run HERE (the container that has
/root/reference); the GPU box only reads the .npz files.

  adabelief_ref.npz   trajectory of the REFERENCE's own optimiser
                      (/root/reference/torch-points3d/torch_points3d/core/optimizer/adabelief.py, pure torch,
                      imported unmodified) on fixed gradients -- pins oracle/train.py:AdaBelief.
  gridsampling_ref.npz  torch expressions of grid_transform.py:116 evaluated with torch CPU (round of the
                      fp32 quotient) -- pins oracle/coords.py:quantize_points against torch semantics.
  msenet14_oracle.npz a small seeded plot batch, its oracle voxelisation / strided maps / kernel-map pair
                      counts and the MSENet14 oracle output -- regression fixture for the oracle itself and a
                      golden input/output pair for the CUDA path (parity unpinned upstream, see oracle/).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/torch-points3d/torch_points3d"


def adabelief_golden():
    spec = importlib.util.spec_from_file_location("ref_adabelief", f"{REF}/core/optimizer/adabelief.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(123)
    p = torch.randn(64, generator=g).requires_grad_()
    p0 = p.detach().clone()
    opt = mod.AdaBelief([p], lr=5e-3, betas=(0.9, 0.999), eps=1e-16, weight_decay=1e-2)
    grads, traj = [], []
    for step in range(12):
        gr = torch.randn(64, generator=g) * (1.0 + step)
        p.grad = gr.clone()
        opt.step()
        grads.append(gr.numpy().copy())
        traj.append(p.detach().numpy().copy())
    np.savez(os.path.join(HERE, "adabelief_ref.npz"), p0=p0.numpy(), grads=np.stack(grads), traj=np.stack(traj))


def gridsampling_golden():
    g = torch.Generator().manual_seed(5)
    pos = torch.rand((4000, 3), generator=g) * torch.tensor([1.0, 1.0, 1.25])
    pos[:64, 0] = (torch.arange(64) + 0.5) * 0.0125          # exact half-way cases
    size = 0.0125
    coords = torch.round(pos / size)                          # grid_transform.py:116, verbatim expression
    np.savez(os.path.join(HERE, "gridsampling_ref.npz"), pos=pos.numpy(), size=np.float32(size),
             coords=coords.numpy())


def msenet_golden():
    from dpcr_agb_b200 import msenet, plots
    from oracle import coords as oc
    from oracle import me_cpu
    batch = plots.synth_batch(5, 0, 2, n_points=1200)
    pos_l = [batch["pos"][batch["batch"] == b] for b in range(2)]
    feat_l = [batch["feats"][batch["batch"] == b] for b in range(2)]
    perm_l = [batch["perm"][:1200], batch["perm"][1200:] - 1200]
    c, f, p, s, cnt = oc.quantize_batch(pos_l, feat_l, 0.05, perm_l)
    levels, cur = {}, c
    for ts in (2, 4, 8, 16):
        cur, _ = oc.stride_map(cur, (ts,) * 3)
        levels[ts] = cur
    stem = oc.kernel_map_table(c, c, 7, (1, 1, 1))
    k3s2 = oc.kernel_map_table(c, levels[2], 3, (1, 1, 1))
    torch.manual_seed(0)
    net = msenet.MSENet(me_cpu, "SENet14", drop_path=0.0)
    net.eval()
    with torch.no_grad():
        y = net(me_cpu.SparseTensor(torch.from_numpy(f), coordinates=torch.from_numpy(c)))
    np.savez_compressed(os.path.join(HERE, "msenet14_oracle.npz"), pos=batch["pos"], feats=batch["feats"],
                        batch=batch["batch"], perm=batch["perm"], coords=c, vox_feats=f, src=s,
                        ts2=levels[2], ts4=levels[4], ts8=levels[8], ts16=levels[16],
                        stem_pairs=(stem >= 0).sum(1), pool_nbr=k3s2, output=y.numpy())


if __name__ == "__main__":
    adabelief_golden()
    gridsampling_golden()
    msenet_golden()
    print("golden fixtures written to", HERE)
