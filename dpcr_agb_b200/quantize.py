"""GPU voxel quantisation -- the ``GridSampling3D(size, quantize_coords=True, mode="last")`` step of the
reference (``torch_points3d/core/data_transform/grid_transform.py:83-135``), batched over plots.

The reference runs this per sample on CPU inside DataLoader workers (``shuffle_data`` :22-29,
``torch.round(pos / size)`` :116, ``grid_cluster`` :117-118, ``consecutive_cluster`` :121, gather of every
per-point tensor at the representative :64-66,123-125).  Here one call handles the whole collated
batch on the GPU through the C ABI (``b2s_quantize_*``): an occupancy bitmap + popcount ranks replace
the sort behind ``torch.unique``; the output row order is identical (sorted by (z, y, x) per plot, plots
concatenated) and the representative of a voxel is the LAST point in the shuffled order.
"""
from __future__ import annotations

import torch

from . import lib as L


class GridSampling3D:
    """Same constructor arguments as the reference class (``mode="mean"`` is not used by any NFI config
    and is not implemented)."""

    def __init__(self, size, quantize_coords=True, mode="last", verbose=False):
        if mode != "last":
            raise NotImplementedError('only mode="last" (all 24 YAML uses in the reference) is implemented')
        self._grid_size = float(size)
        self._quantize_coords = quantize_coords
        self._mode = mode

    def __call__(self, pos, batch=None, tensors=(), order=None, num_plots=None, bounds=None, capacity=None,
                 n_points_dev=None):
        """Quantise a collated batch.

        pos      float32 [n,3] CUDA, positions of all plots concatenated
        batch    int32 [n] plot id per point (None = a single plot)
        tensors  per-point float32 tensors [n,F] to gather at the representative (``data.x`` ...)
        order    int32 [n]: ``order[j]`` = original index of the j-th point of the SHUFFLED sequence
                 (the concatenated per-plot ``torch.randperm`` of ``shuffle_data``); None = identity
        bounds   optional ((lo_x,lo_y,lo_z),(hi_x,hi_y,hi_z)) of the integer grid; when omitted it is
                 measured on the device (one extra host sync)

        capacity optional fixed number of output rows (STATIC mode, needs ``bounds`` and ``num_plots``): outputs
                 are allocated at ``capacity`` rows, the voxel count stays on the device (``num_rows`` int32 [1]) and
                 the call does not synchronise -- the form a captured CUDA graph replays.
        n_points_dev optional int32 [1] device count of live points when ``pos`` is padded to a fixed size

        Returns dict(coords=int32 [M,4] (plot,x,y,z), src=int32 [M], pos=[M,3], tensors=[...], num_rows).
        """
        assert pos.is_cuda and pos.dtype == torch.float32 and pos.dim() == 2 and pos.shape[1] == 3
        pos = pos.contiguous()
        n = pos.shape[0]
        dev = pos.device
        if batch is None:
            batch = torch.zeros(n, dtype=torch.int32, device=dev)
            num_plots = 1
        batch = batch.to(torch.int32).contiguous()
        if num_plots is None:
            num_plots = int(batch.max().item()) + 1 if n else 1
        if order is not None:
            order = order.to(torch.int32).contiguous()

        q = torch.empty((n, 3), dtype=torch.int32, device=dev)
        bnd = torch.empty(6, dtype=torch.int32, device=dev)
        L.call("b2s_quantize_points", pos, n, n_points_dev, self._grid_size, q, bnd)
        static = capacity is not None
        if static:
            assert bounds is not None and num_plots is not None, "static mode needs bounds and num_plots"
        if bounds is None:
            b = bnd.tolist()
            lo, hi = b[:3], b[3:]
        else:
            lo, hi = [int(v) for v in bounds[0]], [int(v) for v in bounds[1]]
        coords, src, m, m_arg, m_dev, index = self._cells(q, batch, order, n, n_points_dev, num_plots, lo, hi, capacity)

        def gather(t):
            t = t.contiguous()
            t2 = t.view(n, -1).float()
            out = torch.empty((m, t2.shape[1]), dtype=torch.float32, device=dev)
            L.call("b2s_gather_rows", t2, src, m, m_arg, t2.shape[1], out)
            return out

        return {"coords": coords, "src": src, "pos": gather(pos), "tensors": [gather(t) for t in tensors],
                "grid_size": self._grid_size, "num_rows": m_dev,
                # the occupancy bitmap + popcount prefix of this batch: rank(cell) == output row, so the coordinate
                # manager can resolve neighbours of THIS map without hash probes (b2s_kernel_map_dense)
                "index": index}

    @staticmethod
    def _cells(q, batch, order, n, n_points_dev, num_plots, lo, hi, capacity):
        """Occupied cells of integer coordinates ``q`` [n,3]: rows sorted by (plot, z, y, x), the representative point
        of every cell (LAST in ``order``) and the occupancy index.  Returns (coords, src, m, m_arg, m_dev, index)."""
        dev = q.device
        static = capacity is not None
        dims = [max(h - l + 1, 1) for l, h in zip(lo, hi)]
        lo_h, dims_h = L.host_i32(*lo), L.host_i32(*dims)
        ws_bytes = L.query("b2s_quantize_workspace_bytes", num_plots, dims_h)
        if ws_bytes < 0:
            raise L.B2SError(f"voxel box {dims} x {num_plots} plots is too large (B2S_EOVERFLOW)")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        m_dev = torch.empty(1, dtype=torch.int32, device=dev)
        L.call("b2s_quantize_count", q, batch, n, n_points_dev, num_plots, lo_h, dims_h, ws, ws_bytes, m_dev)
        if static:
            m, m_arg = int(capacity), m_dev                     # no sync: the count is read on the device
        else:
            m, m_arg = int(m_dev.item()), None                  # host sync: output size
            if m < 0:
                raise L.B2SError("a point falls outside the supplied voxel bounds")
        coords = torch.empty((m, 4), dtype=torch.int32, device=dev)
        src = torch.empty(m, dtype=torch.int32, device=dev)
        L.call("b2s_quantize_fill", q, batch, order, n, n_points_dev, num_plots, lo_h, dims_h, ws, m, m_arg, coords, src)
        return coords, src, m, m_arg, m_dev, (ws, tuple(lo), tuple(dims), int(num_plots))

    def resort(self, vox, num_plots, bounds=None, capacity=None):
        """Rows of ``vox`` (this class's result whose ``coords`` were changed in place by a one-to-one integer map:
        RandomCoordsFlip / ShiftVoxels) back into (plot, z, y, x) order with a fresh occupancy index, every per-row
        tensor permuted along -- what the x-line stem kernels and the dense kernel maps need.  ``bounds`` must hold the
        CHANGED coordinates (measured here when None: one host sync); static mode as in ``__call__``."""
        coords = vox["coords"]
        static = capacity is not None
        n = coords.shape[0]
        n_dev = vox["num_rows"] if static else None
        q = coords[:, 1:4].contiguous()
        plot = coords[:, 0].contiguous()
        if bounds is None:
            assert not static, "static mode needs bounds"
            lo, hi = q.amin(0).tolist(), q.amax(0).tolist()
        else:
            lo, hi = [int(v) for v in bounds[0]], [int(v) for v in bounds[1]]
        new_coords, perm, m, m_arg, m_dev, index = self._cells(q, plot, None, n, n_dev, num_plots, lo, hi, capacity)

        def gather(t, as_int=False):
            t2 = t.contiguous().view(n, -1)
            out = torch.empty((m, t2.shape[1]), dtype=torch.float32, device=t.device)
            L.call("b2s_gather_rows", t2.view(torch.float32) if as_int else t2.float(), perm, m, m_arg, t2.shape[1], out)
            return out.view(torch.int32) if as_int else out

        out = dict(vox)
        out.update(coords=new_coords, src=gather(vox["src"], True).view(-1), pos=gather(vox["pos"]),
                   tensors=[gather(t) for t in vox["tensors"]], num_rows=m_dev, index=index, row_perm=perm)
        return out

    def __repr__(self):
        return "{}(grid_size={}, quantize_coords={}, mode={})".format(
            self.__class__.__name__, self._grid_size, self._quantize_coords, self._mode)
