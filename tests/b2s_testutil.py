"""Shared helpers for the parity tests: seeded synthetic plots, oracle-side quantisation, error norms."""
import numpy as np
import torch

from dpcr_agb_b200 import plots
from oracle import coords as oc

REL_TOL = 1e-3   # north_star: features and gradients within 1e-3 relative (fp32 / TF32)


def make_points(num_plots, n_points, cfg=7, first=0):
    return plots.synth_batch(cfg, first, num_plots, n_points=n_points)


def oracle_quantize(batch, size):
    nb = int(batch["batch"].max()) + 1
    pos_l, feat_l, perm_l = [], [], []
    base = 0
    for b in range(nb):
        sel = batch["batch"] == b
        n = int(sel.sum())
        pos_l.append(batch["pos"][sel])
        feat_l.append(batch["feats"][sel])
        perm_l.append(batch["perm"][base:base + n] - base)
        base += n
    return oc.quantize_batch(pos_l, feat_l, size, perm_l)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max|b|  (b = oracle); the norm SURVEY.md 8c prescribes."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return ((a - b).abs().max() / denom).item()


def assert_close(a, b, tol=REL_TOL, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol:.1e}"
    return e


def random_coords(rng, n, nb=2, extent=24, ts=1):
    """n unique (b,x,y,z) rows on a ts-grid, batch-sorted, otherwise random order."""
    seen = set()
    rows = []
    while len(rows) < n:
        c = (int(rng.integers(nb)), *(int(v) * ts for v in rng.integers(-extent, extent, 3)))
        if c not in seen:
            seen.add(c)
            rows.append(c)
    a = np.asarray(rows, dtype=np.int32)
    return a[np.argsort(a[:, 0], kind="stable")]
