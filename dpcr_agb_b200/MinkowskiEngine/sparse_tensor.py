"""``SparseTensor`` -- the object the reference builds at ``models/instance/minkowski.py:74`` and
re-wraps at ``modules/MinkowskiEngine/common.py:304-308,337-341,363-366,383-386``.

It is deliberately neither a Mapping nor iterable: ``torch.cuda.amp.custom_fwd`` (used by
``senet_block.py:46,126``) rebuilds Mapping / Iterable arguments element-wise, which would strip
the coordinate manager.
"""
from __future__ import annotations

import torch

from .coordinate_manager import CoordinateManager, CoordinateMapKey, _triple


class SparseTensor:
    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager=None, quantization_mode=None, allocator_type=None,
                 minkowski_algorithm=None, requires_grad=None, device=None, num_rows=None, capacities=None,
                 num_batches=None, dense_index=None):
        """Beyond the MinkowskiEngine signature: ``num_rows`` (int32 [1] device tensor), ``capacities`` and
        ``num_batches`` create the tensor in STATIC mode -- ``features`` / ``coordinates`` are allocated at a fixed
        capacity, only the first ``num_rows[0]`` rows are live (unique, batch-sorted), and nothing synchronises.
        ``dense_index``: the ``"index"`` entry of the ``GridSampling3D`` output the coordinates came from; kernel maps
        that look rows up in this tensor's map then use the occupancy index instead of hash probes."""
        assert isinstance(features, torch.Tensor), "features must be a torch.Tensor"
        if device is not None:
            features = features.to(device, non_blocking=True)
        if coordinate_manager is None and num_rows is not None:
            assert coordinates is not None and capacities is not None and num_batches is not None
            coordinate_manager = CoordinateManager(D=3, device=features.device, capacities=capacities,
                                                   num_batches=num_batches)
            coordinate_map_key = coordinate_manager.insert_static(coordinates, num_rows, _triple(tensor_stride),
                                                                  dense_index=dense_index)
        elif coordinate_manager is None:
            assert coordinates is not None, "either coordinates or (coordinate_map_key, coordinate_manager)"
            coordinates = coordinates.to(features.device, non_blocking=True)
            if not features.is_cuda:
                raise RuntimeError("dpcr_agb_b200.MinkowskiEngine runs on CUDA (B200) only: pass device='cuda' -- "
                                   "there is no CPU path")
            coordinate_manager = CoordinateManager(D=coordinates.shape[1] - 1, device=features.device)
            coordinate_map_key, unique_index = coordinate_manager.insert(coordinates, _triple(tensor_stride),
                                                                         dense_index=dense_index)
            if unique_index is not None:
                features = features[unique_index]
        else:
            assert coordinate_map_key is not None
        if requires_grad is not None:
            features.requires_grad_(requires_grad)
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    # ---- the attribute surface the reference touches --------------------------------------
    @property
    def F(self):
        return self._F

    @property
    def features(self):
        return self._F

    @property
    def C(self):
        return self.coordinate_manager.coords(self.coordinate_map_key)

    @property
    def coordinates(self):
        return self.C

    @property
    def n_dev(self):
        """Device row count of this tensor's map (None in dynamic mode)."""
        return self.coordinate_manager.n_dev(self.coordinate_map_key)

    @property
    def tensor_stride(self):
        return list(self.coordinate_map_key.tensor_stride)

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def shape(self):
        return self._F.shape

    @property
    def D(self):
        return self.coordinate_manager.D

    @property
    def requires_grad(self):
        return self._F.requires_grad

    def size(self, *a):
        return self._F.size(*a)

    def dim(self):
        return self._F.dim()

    def __len__(self):
        return self._F.shape[0]

    def _row_slices(self):
        counts = self.coordinate_manager.rows_per_batch(self.coordinate_map_key)
        out, s = [], 0
        for c in counts:
            out.append((s, s + c))
            s += c
        return out

    @property
    def decomposed_coordinates(self):
        """One coordinate tensor per plot, ascending batch id (rows are batch-contiguous in every map this
        manager builds from batch-sorted input; ``common.py:357-359`` relies on exactly that)."""
        c = self.C
        return [c[a:b, 1:] for a, b in self._row_slices()]

    @property
    def decomposed_features(self):
        return [self._F[a:b] for a, b in self._row_slices()]

    def _wrap(self, feats):
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def _same_map(self, other):
        return other.coordinate_manager is self.coordinate_manager and \
            other.coordinate_map_key == self.coordinate_map_key

    def _union_add(self, other, sign=1.0):
        """``a + b`` for tensors on DIFFERENT coordinate maps of one manager: features land on the union map
        (rows of ``a``, then the rows only ``b`` has); every union row receives at most one row of each operand."""
        cm = self.coordinate_manager
        assert other.coordinate_manager is cm, "tensors of different coordinate managers"
        key, ra, rb = cm.union(self.coordinate_map_key, other.coordinate_map_key)
        out = self._F.new_zeros((cm.coords(key).shape[0], self._F.shape[1]))
        out = out.index_add(0, ra, self._F).index_add(0, rb, other._F if sign == 1.0 else -other._F)
        return SparseTensor(out, coordinate_map_key=key, coordinate_manager=cm)

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            if not self._same_map(other):
                return self._union_add(other)
            return self._wrap(self._F + other._F)
        return self._wrap(self._F + other)

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, SparseTensor):
            if not self._same_map(other):
                return self._union_add(other, -1.0)
            return self._wrap(self._F - other._F)
        return self._wrap(self._F - other)

    def __mul__(self, other):
        if isinstance(other, SparseTensor):
            assert self._same_map(other)
            return self._wrap(self._F * other._F)
        return self._wrap(self._F * other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, SparseTensor):
            assert self._same_map(other)
            return self._wrap(self._F / other._F)
        return self._wrap(self._F / other)

    def detach(self):
        return self._wrap(self._F.detach())

    def __repr__(self):
        return (f"SparseTensor(F={tuple(self._F.shape)}, {self.coordinate_map_key}, "
                f"device={self._F.device})")


__all__ = ["SparseTensor", "CoordinateManager", "CoordinateMapKey"]
