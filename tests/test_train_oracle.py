"""Pins the training-step oracle (oracle/train.py) against the REFERENCE's own optimiser: the fixture
tests/golden/adabelief_ref.npz was produced by importing torch_points3d/core/optimizer/adabelief.py unmodified
(tests/golden/make_golden.py); when /root/reference is present the comparison is also made live."""
import importlib.util
import math
import os

import numpy as np
import pytest
import torch

from dpcr_agb_b200 import train
from oracle import train as otrain

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_OPT = "/root/reference/torch-points3d/torch_points3d/core/optimizer/adabelief.py"


def test_adabelief_restatement_matches_reference_fixture():
    d = np.load(os.path.join(GOLD, "adabelief_ref.npz"))
    p = torch.from_numpy(d["p0"].copy()).requires_grad_()
    opt = otrain.AdaBelief([p], lr=5e-3, betas=(0.9, 0.999), eps=1e-16, weight_decay=1e-2)
    for step in range(d["grads"].shape[0]):
        p.grad = torch.from_numpy(d["grads"][step].copy())
        opt.step()
        assert np.allclose(p.detach().numpy(), d["traj"][step], rtol=1e-6, atol=1e-7), step


@pytest.mark.skipif(not os.path.exists(REF_OPT), reason="reference tree not present on this machine")
def test_adabelief_restatement_matches_reference_live():
    spec = importlib.util.spec_from_file_location("ref_adabelief", REF_OPT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(3)
    a = [torch.randn(5, 7).requires_grad_(), torch.randn(9).requires_grad_()]
    b = [t.detach().clone().requires_grad_() for t in a]
    ref = mod.AdaBelief(a, lr=5e-3, weight_decay=1e-2)
    mine = otrain.AdaBelief(b, lr=5e-3, weight_decay=1e-2)
    for _ in range(20):
        for x, y in zip(a, b):
            g = torch.randn_like(x)
            x.grad, y.grad = g.clone(), g.clone()
        ref.step()
        mine.step()
        for x, y in zip(a, b):
            assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)


def test_rectified_step_host_math_agrees():
    for step in (1, 2, 5, 6, 7, 50, 1000):
        assert train.FlatAdaBelief.rectified_step(step, 0.9, 0.999) == otrain.AdaBelief.rectified_step(step, 0.9, 0.999)


def test_cosine_warm_restarts_matches_torch():
    sched = train.CosineAnnealingWarmRestarts(5e-3, T_0=10, T_mult=2)
    lin = torch.nn.Linear(1, 1)
    opt = torch.optim.SGD(lin.parameters(), lr=5e-3)
    ref = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=10, T_mult=2)
    for e in [0.0, 0.0075, 3.3, 9.99, 10.0, 12.5, 29.9, 30.0, 45.2, 71.0, 150.3]:
        ref.step(e)
        assert math.isclose(sched.lr_at(e), opt.param_groups[0]["lr"], rel_tol=1e-9, abs_tol=1e-15)


def test_reg_loss_matches_oracle_and_masks_nan():
    torch.manual_seed(0)
    pred = torch.randn(6, 2)
    tgt = torch.randn(6, 2) * 100 + 100
    c, s = torch.tensor([107.0, 200.0]), torch.tensor([103.0, 194.0])
    assert torch.allclose(train.reg_loss(pred, tgt, c, s), otrain.reg_loss(pred, tgt, c, s))
    tgt[1, 0] = float("nan")
    tgt[4, 1] = float("nan")
    assert torch.allclose(train.reg_loss(pred, tgt, c, s), otrain.reg_loss(pred, tgt, c, s))
