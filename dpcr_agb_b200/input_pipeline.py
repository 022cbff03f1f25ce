"""GPU input pipeline: the per-sample arithmetic transforms of the reference's NFI configuration, batched on the device
in front of the voxel quantiser.

The reference runs them in DataLoader workers on the CPU (``conf/data/instance/NFI/transforms/sparse-xy.yaml:105-152``
and the same steps inside ``train_transform`` :18-104): ScalePos -> MoveCenterPosPerSample -> StartZFromZero ->
Polygon2dExtend (hexagon crop, matplotlib ``contains_points``) -> MaxPoints -> [1, z, distance-to-centre] features ->
GridSampling3D -> RandomCoordsFlip / ShiftVoxels.  At 4 k plots/s per GPU four CPU workers cannot feed the model, so
here the raw collated points (metres, one plot after the other) go to the GPU once and every step is a kernel of
``libb200sparse.so`` (``b2s_plot_transform``, ``b2s_compact_points``, ``b2s_select_by_rank``, ``b2s_point_features``,
``b2s_quantize_*``, ``b2s_coords_augment``).  Random decisions (MaxPoints permutation, GridSampling shuffle, flips,
shifts) are inputs, so the result is reproducible against the CPU oracle bit for bit.
"""
from __future__ import annotations

import torch

from . import lib as L
from .quantize import GridSampling3D

HEXAGON = ((0., 0.5), (0.25, 0.9330127), (0.75, 0.9330127), (1., 0.5), (0.75, 0.0669873), (0.25, 0.0669873))


class NFIInputPipeline:
    def __init__(self, scale=(30.0, 30.0, 40.0), center=(0.5, 0.5), polygon=HEXAGON, max_points=16000,
                 grid_size=0.0125):
        self.scale = tuple(float(v) for v in scale)
        self.center = (float(center[0]), float(center[1]), 0.5)      # MoveCenterPosPerSample's default center_z
        self.polygon = tuple((float(x), float(y)) for x, y in polygon) if polygon else ()
        assert len(self.polygon) <= 16
        self.max_points = int(max_points)
        self.gs = GridSampling3D(grid_size)

    # ---- steps ------------------------------------------------------------------------------------------------
    def transform(self, raw_pos, plot, num_plots, n_dev=None):
        """ScalePos + MoveCenterPosPerSample + StartZFromZero + the polygon test.  Returns (pos, keep int32 [n])."""
        raw_pos = raw_pos.contiguous()
        n = raw_pos.shape[0]
        dev = raw_pos.device
        pos, keep = torch.empty_like(raw_pos), torch.empty(n, dtype=torch.int32, device=dev)
        minz = torch.empty(num_plots, dtype=torch.int32, device=dev)
        flat = [v for xy in self.polygon for v in xy]
        L.call("b2s_plot_transform", raw_pos, plot, n, n_dev, num_plots, L.host_f32(*self.scale),
               L.host_f32(*self.center), L.host_f64(*flat) if flat else None, len(self.polygon), minz, pos, keep)
        return pos, keep

    def compact(self, pos, plot, keep, n_dev=None):
        """Polygon2dExtend's ``apply_mask``: surviving points in order.  Returns (pos, plot, count int32 [1] device);
        the arrays keep their capacity, only the first ``count`` rows are live."""
        n, dev = pos.shape[0], pos.device
        out_pos, out_plot = torch.empty_like(pos), torch.empty_like(plot)
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        ws = torch.empty(L.query("b2s_scan_workspace_bytes", n), dtype=torch.uint8, device=dev)
        count = torch.empty(1, dtype=torch.int32, device=dev)
        L.call("b2s_compact_points", pos, plot, keep, n, n_dev, idx, ws, out_pos, out_plot, count)
        return out_pos, out_plot, count

    def max_points_select(self, pos, plot, num_plots, rank, n_dev=None):
        """MaxPoints with the permutation given as ``rank`` (int32 [n]: position of every point in its plot's
        ``randperm``).  Returns (pos, plot, count int32 [1])."""
        n, dev = pos.shape[0], pos.device
        counts = torch.empty(num_plots, dtype=torch.int32, device=dev)
        L.call("b2s_batch_counts", plot, 1, n, n_dev, num_plots, counts)
        offsets = torch.empty(num_plots + 1, dtype=torch.int32, device=dev)
        out_pos, out_plot = torch.empty_like(pos), torch.empty_like(plot)
        L.call("b2s_select_by_rank", pos, plot, rank, n, n_dev, num_plots, self.max_points, counts, offsets, out_pos,
               out_plot)
        return out_pos, out_plot, offsets[num_plots:num_plots + 1]

    def features(self, pos, n_dev=None):
        feats = torch.empty((pos.shape[0], 3), dtype=torch.float32, device=pos.device)
        L.call("b2s_point_features", pos, pos.shape[0], n_dev, self.center[0], self.center[1], feats)
        return feats

    def augment_coords(self, coords, num_plots, flips, shifts, m_dev=None):
        """RandomCoordsFlip(ignored z) + ShiftVoxels in place: ``flips`` [B,2] / ``shifts`` [B,3] integer tensors."""
        aug = torch.cat([torch.as_tensor(flips).reshape(num_plots, 2), torch.as_tensor(shifts).reshape(num_plots, 3)],
                        1).to(device=coords.device, dtype=torch.int32).contiguous()
        scratch = torch.empty((num_plots, 2), dtype=torch.int32, device=coords.device)
        L.call("b2s_coords_augment", coords, coords.shape[0], m_dev, num_plots, aug, scratch)
        return coords

    # ---- the whole chain --------------------------------------------------------------------------------------
    def __call__(self, raw_pos, plot, num_plots, order=None, max_points_rank=None, flips=None, shifts=None,
                 bounds=None, capacity=None, n_points_dev=None, resort=True, aug_bounds=None):
        """raw_pos float32 [n,3] in metres, plot int32 [n] (plots contiguous).  ``order``: GridSampling3D's shuffle
        over the SURVIVING points (see :class:`GridSampling3D`); ``max_points_rank``: enables MaxPoints.  With
        ``bounds`` + ``capacity`` nothing synchronises with the host (captured-graph form).  Returns the quantiser's
        dict plus ``num_points`` (int32 [1], points fed to the quantiser)."""
        plot = plot.to(torch.int32).contiguous()
        pos, keep = self.transform(raw_pos, plot, num_plots, n_points_dev)
        pos, plot, count = self.compact(pos, plot, keep, n_points_dev)
        if max_points_rank is not None:
            pos, plot, count = self.max_points_select(pos, plot, num_plots, max_points_rank.to(torch.int32), count)
        feats = self.features(pos, count)
        static = capacity is not None
        if not static:                       # dynamic mode: one host sync for the surviving point count
            n = int(count.item())
            pos, plot, feats, count_arg = pos[:n].contiguous(), plot[:n].contiguous(), feats[:n].contiguous(), None
        else:
            count_arg = count
        vox = self.gs(pos, plot, tensors=(feats,), order=order, num_plots=num_plots, bounds=bounds, capacity=capacity,
                      n_points_dev=count_arg)
        if flips is not None or shifts is not None:
            flips = flips if flips is not None else torch.zeros((num_plots, 2), dtype=torch.int32)
            shifts = shifts if shifts is not None else torch.zeros((num_plots, 3), dtype=torch.int32)
            self.augment_coords(vox["coords"], num_plots, flips, shifts, vox["num_rows"] if static else None)
            vox["index"] = None              # the occupancy index describes the un-augmented coordinates
            # The reference leaves the rows in the quantiser's order with the changed coordinates (the network is
            # indifferent to the row order).  ``resort``: bring them back into (plot, z, y, x) order with a fresh
            # occupancy index, which keeps the x-line stem kernels and the dense kernel maps on the augmented batch
            # (otherwise the hash-probed [343, N] table path: +1.1 ms per 32-plot step).  Static mode needs the box of
            # the augmented coordinates (``aug_bounds``); without it the rows stay as they are.
            if resort and (not static or aug_bounds is not None):
                vox = self.gs.resort(vox, num_plots, aug_bounds, capacity)
        vox["num_points"] = count
        return vox
