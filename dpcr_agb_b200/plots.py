"""Synthetic "Danish-NFI-shaped" LiDAR plots -- the benchmark input.

The real data (LAS files, ``README.md:119-126`` of the reference) cannot be downloaded here, so plots
are generated to the same contract the reference's transform list produces
(``conf/data/instance/NFI/transforms/sparse-xy.yaml:105-152``): positions divided by (30, 30, 40) and
shifted by +0.5 in x,y, z starting at 0, cropped to the unit hexagon (:117-123), at most 16 000 points
(``MaxPoints`` :124-127), features ``[ones, pos_z, xy_distance]`` (:132-147), to be voxelised with
``GridSampling3D(size=0.0125)`` (:148-152; ``conf/data/instance/NFI/default.yaml:23``).
Canopy heights follow the ``h_q99`` statistics of ``nfi-data/train_split.csv`` (mean 16.5 m, sd 7.8 m).
"""
from __future__ import annotations

import numpy as np

HEXAGON = np.array([[0.0, 0.5], [0.25, 0.9330127], [0.75, 0.9330127],
                    [1.0, 0.5], [0.75, 0.0669873], [0.25, 0.0669873]], dtype=np.float64)
GRID_SIZE = 0.0125
SCALE_XYZ = (30.0, 30.0, 40.0)
TARGET_MEAN = np.array([107.0, 200.0], dtype=np.float32)   # BMag_ha, V_ha of train_split.csv
TARGET_STD = np.array([103.0, 194.0], dtype=np.float32)


def _in_hexagon(x, y):
    """Point-in-convex-polygon test for the (counter-clockwise re-ordered) unit hexagon."""
    v = HEXAGON[::-1]                      # the YAML lists the vertices clockwise
    inside = np.ones(x.shape, dtype=bool)
    for i in range(6):
        x0, y0 = v[i]
        x1, y1 = v[(i + 1) % 6]
        inside &= (x1 - x0) * (y - y0) - (y1 - y0) * (x - x0) >= 0.0
    return inside


def synth_plot(seed: int, n_points: int = 16000, canopy_max_m: float = 40.0):
    """One plot: returns (pos float32 [n,3] normalised, feats float32 [n,3], target float32 [2])."""
    rng = np.random.default_rng(seed)
    h_top = float(np.clip(rng.normal(16.5, 7.8), 2.0, canopy_max_m))
    xs, ys = [], []
    need = n_points
    while need > 0:
        x = rng.random(2 * need + 64)
        y = rng.random(2 * need + 64)
        keep = _in_hexagon(x, y)
        xs.append(x[keep][:need])
        ys.append(y[keep][:need])
        need -= xs[-1].shape[0]
    x = np.concatenate(xs)
    y = np.concatenate(ys)
    ground = rng.random(n_points) < 0.3
    z_m = np.where(ground, np.abs(rng.normal(0.0, 0.05, n_points)), rng.beta(4.0, 2.0, n_points) * h_top)
    z_m = np.clip(z_m, 0.0, 50.0)
    z = z_m / SCALE_XYZ[2]
    z = z - z.min()                                               # StartZFromZero
    pos = np.stack([x, y, z], 1).astype(np.float32)
    dist = np.sqrt((pos[:, 0] - 0.5) ** 2 + (pos[:, 1] - 0.5) ** 2)
    feats = np.stack([np.ones(n_points, np.float32), pos[:, 2], dist.astype(np.float32)], 1).astype(np.float32)
    frac = h_top / 16.5
    target = (TARGET_MEAN * frac * (1.0 + 0.1 * rng.standard_normal(2))).astype(np.float32)
    return pos, feats, target


def synth_batch(cfg: int, first_plot: int, num_plots: int, n_points: int = 16000, canopy_max_m: float = 40.0):
    """A collated batch: seed of plot p is ``1000*cfg + p``.  Returns dict of numpy arrays:
    pos [sum n,3], feats [sum n,3], batch int32 [sum n], target [B,2], perm int32 [sum n] (per-plot shuffles
    already offset into the concatenated array -- the explicit ``shuffle_data`` permutation)."""
    pos, feats, batch, target, perm = [], [], [], [], []
    base = 0
    for b in range(num_plots):
        seed = 1000 * cfg + first_plot + b
        p, f, t = synth_plot(seed, n_points, canopy_max_m)
        pos.append(p)
        feats.append(f)
        batch.append(np.full(p.shape[0], b, np.int32))
        target.append(t)
        perm.append(np.random.default_rng(seed + 500_000).permutation(p.shape[0]).astype(np.int32) + base)
        base += p.shape[0]
    return {"pos": np.concatenate(pos), "feats": np.concatenate(feats), "batch": np.concatenate(batch),
            "target": np.stack(target), "perm": np.concatenate(perm)}


def synth_batch_from_ids(cfg: int, plot_ids, n_points: int = 16000, canopy_max_m: float = 40.0):
    """Like :func:`synth_batch` for an arbitrary list of plot ids (seed of plot p is ``1000*cfg + p``): the batch a
    data-parallel rank assembles from the plots ``train.shard_plots`` dealt to it."""
    pos, feats, batch, target, perm = [], [], [], [], []
    base = 0
    for b, pid in enumerate(plot_ids):
        seed = 1000 * cfg + int(pid)
        p, f, t = synth_plot(seed, n_points, canopy_max_m)
        pos.append(p)
        feats.append(f)
        batch.append(np.full(p.shape[0], b, np.int32))
        target.append(t)
        perm.append(np.random.default_rng(seed + 500_000).permutation(p.shape[0]).astype(np.int32) + base)
        base += p.shape[0]
    return {"pos": np.concatenate(pos), "feats": np.concatenate(feats), "batch": np.concatenate(batch),
            "target": np.stack(target), "perm": np.concatenate(perm)}


def voxel_count(pos: np.ndarray, grid: float) -> int:
    """Number of occupied voxels of one plot at ``grid`` (the work estimate ``train.shard_plots`` balances ranks by)."""
    q = np.rint(pos / np.float32(grid)).astype(np.int64)
    return int(np.unique((q[:, 2] * 4096 + q[:, 1]) * 4096 + q[:, 0]).shape[0])
