"""Training-step half of the oracle: regression loss and AdaBelief, restated on CPU.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Unlike the sparse ops, these two pieces live in
the reference tree itself, so this restatement IS pinned: ``tests/golden/make_adabelief_golden.py``
imports ``torch_points3d/core/optimizer/adabelief.py`` (pure torch, importable here) and stores its
trajectory as a fixture that ``tests/test_train_oracle.py`` replays against :class:`AdaBelief`.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def reg_loss(pred, target, center, scale, weight=0.5):
    """``InstanceBase.compute_reg_loss`` -- models/instance/base.py:154-179 with the README's
    smooth-L1 loss and ``reg_weights = [0.5, 0.5]`` (conf/data/instance/NFI/reg.yaml:21-24): the targets are
    z-scored, NaN targets are masked out, ``loss = mean(reg_weights) * smooth_l1(pred, labels)``."""
    labels = (target - center) / scale
    mask = ~torch.isnan(labels)
    if not bool(mask.all()):
        pred, labels = pred[mask], labels[mask]
    return weight * F.smooth_l1_loss(pred, labels)


class AdaBelief:
    """core/optimizer/adabelief.py:90-201 with the reference's defaults used by
    conf/training/nfi/minkowski.yaml:9-20 (decoupled decay, rectify, degenerated_to_sgd, no amsgrad)."""

    def __init__(self, params, lr=5e-3, betas=(0.9, 0.999), eps=1e-16, weight_decay=1e-2):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.step_count = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.s = [torch.zeros_like(p) for p in self.params]

    @staticmethod
    def rectified_step(step, beta1, beta2):
        """adabelief.py:169-187 -> (num_sma, step_size)."""
        beta2_t = beta2 ** step
        sma_max = 2.0 / (1.0 - beta2) - 1.0
        sma = sma_max - 2.0 * step * beta2_t / (1.0 - beta2_t)
        if sma >= 5:
            step_size = math.sqrt((1 - beta2_t) * (sma - 4) / (sma_max - 4) * (sma - 2) / sma
                                  * sma_max / (sma_max - 2)) / (1 - beta1 ** step)
        else:
            step_size = 1.0 / (1 - beta1 ** step)      # degenerated_to_sgd=True
        return sma, step_size

    @torch.no_grad()
    def step(self):
        b1, b2 = self.betas
        self.step_count += 1
        sma, step_size = self.rectified_step(self.step_count, b1, b2)
        for p, m, s in zip(self.params, self.m, self.s):
            if p.grad is None:
                continue
            g = p.grad
            p.mul_(1.0 - self.lr * self.wd)                       # :131-135
            m.mul_(b1).add_(g, alpha=1 - b1)                      # :147
            r = g - m
            s.mul_(b2).addcmul_(r, r, value=1 - b2)               # :149
            s.add_(self.eps)                                      # :159 (in place, as upstream)
            if sma >= 5:
                p.addcdiv_(m, s.sqrt().add_(self.eps), value=-step_size * self.lr)   # :193-195
            elif step_size > 0:
                p.add_(m, alpha=-step_size * self.lr)             # :196-197


def cpu_training_step(model, opt, batch, size, center, scale):
    """One whole optimisation step of the path on the CPU oracle, from raw points: per-plot GridSampling3D,
    collate + batch column (minkowski.py:69), SparseTensor (hash / maps built lazily by the ops), forward,
    loss, backward, AdaBelief.  This is what bench.py times as the CPU baseline ("port": a restatement of the
    reference's CPU MinkowskiEngine path, NOT upstream ME-CPU itself)."""
    import numpy as np

    from . import coords as oc
    from . import me_cpu

    nb = int(batch["batch"].max()) + 1
    pos_l, feat_l, perm_l, base = [], [], [], 0
    for b in range(nb):
        sel = batch["batch"] == b
        n = int(sel.sum())
        pos_l.append(batch["pos"][sel])
        feat_l.append(batch["feats"][sel])
        perm_l.append(batch["perm"][base:base + n] - base)
        base += n
    c, f, _, _, _ = oc.quantize_batch(pos_l, feat_l, size, perm_l)
    x = me_cpu.SparseTensor(torch.from_numpy(np.ascontiguousarray(f)), coordinates=torch.from_numpy(c))
    model.train()
    for p in model.parameters():
        p.grad = None
    pred = model(x)
    loss = reg_loss(pred, torch.from_numpy(batch["target"]), center, scale)
    loss.backward()
    for p in model.parameters():                      # clip_grad_value_(100), base_model.py:243
        if p.grad is not None:
            p.grad.clamp_(-100.0, 100.0)
    opt.step()
    return float(loss)


def cpu_inference_step(model, batch, size):
    """Eval forward of a batch on the CPU oracle from raw points (BASELINE.json configs[0]: the reference's CPU
    MinkowskiEngine case, ``eval.py`` -> ``model.forward`` under no_grad, models/base_model.py:153-160)."""
    import numpy as np

    from . import coords as oc
    from . import me_cpu

    nb = int(batch["batch"].max()) + 1
    pos_l, feat_l, perm_l, base = [], [], [], 0
    for b in range(nb):
        sel = batch["batch"] == b
        n = int(sel.sum())
        pos_l.append(batch["pos"][sel])
        feat_l.append(batch["feats"][sel])
        perm_l.append(batch["perm"][base:base + n] - base)
        base += n
    c, f, _, _, _ = oc.quantize_batch(pos_l, feat_l, size, perm_l)
    x = me_cpu.SparseTensor(torch.from_numpy(np.ascontiguousarray(f)), coordinates=torch.from_numpy(c))
    model.eval()
    with torch.no_grad():
        return model(x)
