"""CPU restatement of the reference's per-sample input transforms on the MSENet path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference classes live in
``torch_points3d/core/data_transform/{transforms,features,sparse_transforms}.py``; that package cannot be imported
here (it pulls in torch_geometric, torch_cluster, matplotlib, numba at import time), so each step below restates the
few lines of its ``__call__`` with the same torch / numpy expressions, citing them.  ``Path.contains_points`` of
matplotlib (absent) is restated from its documented crossings algorithm -- PARITY UNPINNED for that one step.
"""
from __future__ import annotations

import numpy as np
import torch

HEXAGON = [[0., 0.5], [0.25, 0.9330127], [0.75, 0.9330127], [1., 0.5], [0.75, 0.0669873], [0.25, 0.0669873]]


def scale_pos(pos: torch.Tensor, scale) -> torch.Tensor:
    """``ScalePos(op="div")``: ``data.pos = torch.div(data.pos, scale)`` (transforms.py:590-598)."""
    return torch.div(pos, torch.tensor(scale, dtype=torch.float32).unsqueeze(0))


def move_center(pos: torch.Tensor, center_x=0.5, center_y=0.5, center_z=0.5) -> torch.Tensor:
    """``MoveCenterPosPerSample``: ``data.pos += center_`` (transforms.py:734-739; center_z defaults to 0.5)."""
    return pos + torch.FloatTensor([[center_x, center_y, center_z]])


def start_z_from_zero(pos: torch.Tensor) -> torch.Tensor:
    """``StartZFromZero``: ``data.pos[:, 2] -= data.pos[:, 2].min()`` (transforms.py:766-769)."""
    pos = pos.clone()
    pos[:, 2] -= pos[:, 2].min()
    return pos


def contains_points(polygon, xy: np.ndarray) -> np.ndarray:
    """``matplotlib.path.Path(polygon).contains_points(xy)`` (radius 0) -- the crossings test of its
    ``point_in_path_impl``, in float64 as matplotlib converts its inputs."""
    v = np.asarray(polygon, dtype=np.float64)
    tx, ty = xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64)
    inside = np.zeros(xy.shape[0], dtype=bool)
    x0, y0 = v[-1]
    yflag0 = y0 >= ty
    for x1, y1 in v:
        yflag1 = y1 >= ty
        cross = ((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == yflag1
        inside ^= (yflag0 != yflag1) & cross
        yflag0, x0, y0 = yflag1, x1, y1
    return inside


def polygon_extend(pos: torch.Tensor, polygon=HEXAGON):
    """``Polygon2dExtend``: ``mask = polygon.contains_points(pos[:, [0, 1]])``; ``apply_mask`` keeps the rows in order
    (transforms.py:1489-1496, 1090-1095).  Returns (pos[mask], mask)."""
    mask = torch.from_numpy(contains_points(polygon, pos[:, [0, 1]].numpy()))
    return pos[mask], mask


def max_points(pos: torch.Tensor, num: int, perm: torch.Tensor | None):
    """``MaxPoints``: if more than ``num`` points, ``choice = torch.randperm(num_nodes)[:num]`` and every per-point
    tensor becomes ``item[choice]`` (transforms.py:1337-1358, 1783-1791).  ``perm`` = that randperm."""
    if pos.shape[0] <= num:
        return pos
    if callable(perm):
        perm = perm(pos.shape[0])
    return pos[perm[:num]]


def features(pos: torch.Tensor, center_x=0.5, center_y=0.5) -> torch.Tensor:
    """``AddOnes`` + ``XYZFeature(add_z)`` + ``AddXYDistanceToCenter`` + ``AddFeatsByKeys([ones, pos_z, xy_distance])``
    (features.py:307-383): x = [1, z, PairwiseDistance()(pos[:, :2], centre)]."""
    center = torch.tensor([[center_x, center_y]])
    d = torch.nn.PairwiseDistance()(pos[:, :2], center.repeat_interleave(pos.shape[0], dim=0))
    return torch.stack([torch.ones(pos.shape[0]), pos[:, 2].clone(), d], 1)


def test_transform(raw_pos: torch.Tensor, scale=(30.0, 30.0, 40.0), center=(0.5, 0.5), polygon=HEXAGON,
                   num=16000, perm=None):
    """``sparse_xy.test_transform`` up to the quantiser (sparse-xy.yaml:105-147) for ONE plot: (pos, x)."""
    pos = scale_pos(raw_pos, scale)
    pos = move_center(pos, center[0], center[1])
    pos = start_z_from_zero(pos)
    pos, _ = polygon_extend(pos, polygon)
    pos = max_points(pos, num, perm)
    return pos, features(pos, center[0], center[1])


def coords_flip(coords: np.ndarray, flip_x: bool, flip_y: bool) -> np.ndarray:
    """``RandomCoordsFlip(ignored_axis="z")`` for one sample (transforms.py:1046-1054): per flipped axis
    ``coords[:, ax] = coords[:, ax].max() - coords[:, ax]``."""
    c = coords.copy()
    for ax, f in ((0, flip_x), (1, flip_y)):
        if f:
            c[:, ax] = c[:, ax].max() - c[:, ax]
    return c


def shift_voxels(coords: np.ndarray, shift) -> np.ndarray:
    """``ShiftVoxels`` (sparse_transforms.py:49-55): ``coords[:, :3] += (torch.rand(3) * 100).type_as(coords)``."""
    return coords + np.asarray(shift, dtype=coords.dtype)[None, :]
