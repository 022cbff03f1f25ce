"""A MinkowskiEngine-shaped namespace over the CPU oracle.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED.  It exists so that the
UNCHANGED reference networks (``torch_points3d/modules/MinkowskiEngine/SENet.py`` etc.) and our own
restatement of them can be executed on CPU as the checker for the CUDA path.  It exports exactly the
``ME.*`` names those files touch -- nothing more.

``install()`` registers this module as ``MinkowskiEngine`` in ``sys.modules`` (tests only).
"""
from __future__ import annotations

import sys
import types
from enum import Enum

import numpy as np
import torch
import torch.nn as nn

from . import coords as oc
from . import ops as oo

__version__ = "oracle-cpu"


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return tuple(int(a) for a in v)
    if isinstance(v, torch.Tensor) or isinstance(v, np.ndarray):
        return tuple(int(a) for a in v)
    return (int(v),) * 3


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class CoordinateMapKey:
    def __init__(self, tensor_stride, tag=""):
        self.tensor_stride = tuple(tensor_stride)
        self.tag = tag

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def get_key(self):
        return (list(self.tensor_stride), self.tag)

    def _k(self):
        return (self.tensor_stride, self.tag)

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and self._k() == other._k()

    def __hash__(self):
        return hash(self._k())

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, tag={self.tag!r})"


class CoordinateManager:
    """Owns coordinate arrays per key and a cache of neighbour tables (numpy, host)."""

    def __init__(self, D=3):
        self.D = D
        self.maps = {}
        self.kmaps = {}
        self.batch_info = {}

    def insert(self, coords: np.ndarray, tensor_stride=(1, 1, 1), tag=""):
        key = CoordinateMapKey(tensor_stride, tag)
        self.maps[key] = np.ascontiguousarray(coords, dtype=np.int32)
        return key

    def coords(self, key):
        return self.maps[key]

    def stride(self, key, stride):
        ts = tuple(a * b for a, b in zip(key.tensor_stride, stride))
        out_key = CoordinateMapKey(ts, key.tag)
        if out_key not in self.maps:
            self.maps[out_key], _ = oc.stride_map(self.maps[key], ts)
        return out_key

    def origin(self, key):
        okey = CoordinateMapKey((0, 0, 0), "origin")
        if okey not in self.maps:
            b = self.maps[key][:, 0]
            nb = int(b.max()) + 1 if b.size else 0
            oc_ = np.zeros((nb, 4), np.int32)
            oc_[:, 0] = np.arange(nb)
            self.maps[okey] = oc_
        return okey

    def num_batches(self):
        k = next(iter(self.maps))
        b = self.maps[k][:, 0]
        return int(b.max()) + 1 if b.size else 0

    def batch_of(self, key):
        if key not in self.batch_info:
            self.batch_info[key] = self.maps[key][:, 0].astype(np.int64)
        return self.batch_info[key]

    def kernel_map(self, in_key, out_key, kernel_size, dilation):
        ck = (in_key, out_key, kernel_size, dilation)
        if ck not in self.kmaps:
            step = tuple(d * t for d, t in zip(dilation, in_key.tensor_stride))
            self.kmaps[ck] = oc.kernel_map_table(self.maps[in_key], self.maps[out_key], kernel_size, step)
        return self.kmaps[ck]

    def transposed_kernel_map(self, coarse_key, fine_key, kernel_size, dilation):
        """Table of a transposed convolution / pooling from ``coarse_key`` rows to ``fine_key`` rows:
        ``nbr[k, f] = c`` iff ``fine[f] == coarse[c] + delta_k`` with the offsets of the FORWARD op fine -> coarse
        (step = dilation * fine tensor stride) -- MinkowskiEngine builds the forward kernel map and swaps its sides."""
        ck = ("T", coarse_key, fine_key, kernel_size, dilation)
        if ck not in self.kmaps:
            step = tuple(d * t for d, t in zip(dilation, fine_key.tensor_stride))
            self.kmaps[ck] = oc.kernel_map_table(self.maps[coarse_key], self.maps[fine_key], kernel_size, step, sign=-1)
        return self.kmaps[ck]

    def union(self, key_a, key_b):
        """Union coordinate map of two maps of the same tensor stride (``SparseTensor.__add__`` across maps): rows of
        ``key_a`` in their order, then the rows only ``key_b`` has, in its order.  Returns (key, rows of a, rows of b)."""
        assert key_a.tensor_stride == key_b.tensor_stride
        both = np.concatenate([self.maps[key_a], self.maps[key_b]])
        first, inv = oc.unique_first(both)
        key = CoordinateMapKey(key_a.tensor_stride, f"union({key_a.tag}|{key_b.tag}|{len(self.maps)})")
        self.maps[key] = both[first]
        na = self.maps[key_a].shape[0]
        return key, inv[:na].astype(np.int64), inv[na:].astype(np.int64)


class SparseTensor:
    """``ME.SparseTensor`` as the reference uses it (``models/instance/minkowski.py:74``,
    ``modules/MinkowskiEngine/common.py:304-308``).  Deliberately neither a Mapping nor iterable
    (``custom_fwd`` in ``senet_block.py:46`` rebuilds those element-wise)."""

    def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None,
                 tensor_stride=1, device=None, **kw):
        if device is not None:
            features = features.to(device)
        if coordinate_manager is None:
            assert coordinates is not None
            c = coordinates.detach().cpu().numpy().astype(np.int32)
            first, inv = oc.unique_first(c)
            if first.shape[0] != c.shape[0]:                  # duplicate rows: keep first occurrence
                c = c[first]
                features = features[torch.from_numpy(first)]
            coordinate_manager = CoordinateManager(D=c.shape[1] - 1)
            coordinate_map_key = coordinate_manager.insert(c, _triple(tensor_stride))
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    @property
    def F(self):
        return self._F

    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.coords(self.coordinate_map_key))

    coordinates = C
    features = F

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

    def size(self, *a):
        return self._F.size(*a)

    def __len__(self):
        return self._F.shape[0]

    @property
    def decomposed_coordinates(self):
        c = self.C
        nb = self.coordinate_manager.num_batches()
        return [c[c[:, 0] == b, 1:] for b in range(nb)]

    @property
    def decomposed_features(self):
        b = torch.from_numpy(self.coordinate_manager.batch_of(self.coordinate_map_key))
        return [self._F[b == i] for i in range(self.coordinate_manager.num_batches())]

    def _wrap(self, f):
        return SparseTensor(f, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self.coordinate_manager)

    def _union_add(self, other, sign=1.0):
        cm = self.coordinate_manager
        key, ra, rb = cm.union(self.coordinate_map_key, other.coordinate_map_key)
        out = self._F.new_zeros((cm.coords(key).shape[0], self._F.shape[1]))
        out = out.index_add(0, torch.from_numpy(ra), self._F).index_add(0, torch.from_numpy(rb), sign * other._F)
        return SparseTensor(out, coordinate_map_key=key, coordinate_manager=cm)

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            assert other.coordinate_manager is self.coordinate_manager
            if other.coordinate_map_key != self.coordinate_map_key:
                return self._union_add(other)
            return self._wrap(self._F + other._F)
        return self._wrap(self._F + other)

    def __mul__(self, other):
        if isinstance(other, SparseTensor):
            assert other.coordinate_map_key == self.coordinate_map_key
            return self._wrap(self._F * other._F)
        return self._wrap(self._F * other)

    def __repr__(self):
        return f"SparseTensor(F={tuple(self._F.shape)}, key={self.coordinate_map_key})"


class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False,
                 region_type=RegionType.HYPER_CUBE, region_offsets=None, expand_coordinates=False,
                 axis_types=None, dimension=-1):
        self.kernel_size = _triple(kernel_size)
        self.kernel_stride = _triple(stride)
        self.kernel_dilation = _triple(dilation)
        self.region_type = region_type
        self.kernel_volume = int(np.prod(self.kernel_size))
        self.dimension = dimension


class MinkowskiModuleBase(nn.Module):
    pass


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class MinkowskiConvolution(MinkowskiModuleBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3, "oracle restates the D=3 path only"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)
        self.kernel_volume = int(np.prod(self.kernel_size))
        self.use_mm = self.kernel_volume == 1 and self.stride == (1, 1, 1)
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            std = 1.0 / np.sqrt(self.in_channels * self.kernel_volume)
            self.kernel.uniform_(-std, std)
            if self.bias is not None:
                self.bias.uniform_(-std, std)

    def forward(self, x: SparseTensor):
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        if self.use_mm:
            return SparseTensor(oo.conv(x.F, self.kernel, None, self.bias), coordinate_map_key=in_key,
                                coordinate_manager=cm)
        out_key = in_key if self.stride == (1, 1, 1) else cm.stride(in_key, self.stride)
        nbr = cm.kernel_map(in_key, out_key, self.kernel_size, self.dilation)
        return SparseTensor(oo.conv(x.F, self.kernel, nbr, self.bias), coordinate_map_key=out_key,
                            coordinate_manager=cm)


class MinkowskiConvolutionTranspose(MinkowskiModuleBase):
    """Non-generative transposed convolution (``modules/MinkowskiEngine/networks.py:155-176``): the output lives on
    the EXISTING map of tensor stride ``ts_in / stride`` (the encoder map of a U-Net);
    ``out[f] = bias + sum_k in[c] @ W[k]`` over the pairs (f, c, k) of the forward op fine -> coarse, sides swapped."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3 and not expand_coordinates
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)
        self.kernel_volume = int(np.prod(self.kernel_size))
        self.kernel = nn.Parameter(torch.empty(self.kernel_volume, in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        with torch.no_grad():
            std = 1.0 / np.sqrt(out_channels * self.kernel_volume)
            self.kernel.uniform_(-std, std)
            if self.bias is not None:
                self.bias.uniform_(-std, std)

    def forward(self, x):
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        ts = tuple(t // s for t, s in zip(in_key.tensor_stride, self.stride))
        out_key = CoordinateMapKey(ts, in_key.tag)
        assert out_key in cm.maps, "oracle: transposed convolution onto an existing (encoder) map only"
        nbr = cm.transposed_kernel_map(in_key, out_key, self.kernel_size, self.dilation)
        return SparseTensor(oo.conv(x.F, self.kernel, nbr, self.bias), coordinate_map_key=out_key,
                            coordinate_manager=cm)


class MinkowskiMaxPooling(MinkowskiModuleBase):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)

    def forward(self, x):
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        out_key = in_key if self.stride == (1, 1, 1) else cm.stride(in_key, self.stride)
        nbr = cm.kernel_map(in_key, out_key, self.kernel_size, self.dilation)
        return SparseTensor(oo.max_pool(x.F, nbr), coordinate_map_key=out_key, coordinate_manager=cm)


class _NotOnPath(MinkowskiModuleBase):
    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, *a, **kw):
        raise NotImplementedError(f"oracle: {type(self).__name__} is outside the MSENet hot path")


class MinkowskiSumPooling(MinkowskiModuleBase):
    """Local sum pooling; ``AVERAGE`` divides by the number of inputs under the kernel (MinkowskiAvgPooling,
    ``networks.py:29``)."""
    AVERAGE = False

    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)

    def forward(self, x):
        cm, in_key = x.coordinate_manager, x.coordinate_map_key
        out_key = in_key if self.stride == (1, 1, 1) else cm.stride(in_key, self.stride)
        nbr = torch.from_numpy(cm.kernel_map(in_key, out_key, self.kernel_size, self.dilation).astype(np.int64))
        out = x.F.new_zeros((nbr.shape[1], x.F.shape[1]))
        for k in range(nbr.shape[0]):
            o = torch.nonzero(nbr[k] >= 0).squeeze(1)
            out = out.index_add(0, o, x.F[nbr[k, o]])
        if self.AVERAGE:
            out = out / (nbr >= 0).sum(0).clamp(min=1).to(out.dtype)[:, None]
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiAvgPooling(MinkowskiSumPooling):
    AVERAGE = True


class MinkowskiAvgUnpooling(_NotOnPath):
    pass


class MinkowskiPoolingTranspose(_NotOnPath):
    pass


class _GlobalPool(MinkowskiModuleBase):
    MODE = "avg"

    def __init__(self, mode=None):
        super().__init__()

    def forward(self, x):
        cm = x.coordinate_manager
        okey = cm.origin(x.coordinate_map_key)
        out = oo.global_pool(x.F, cm.batch_of(x.coordinate_map_key), cm.num_batches(), self.MODE)
        return SparseTensor(out, coordinate_map_key=okey, coordinate_manager=cm)


class MinkowskiGlobalPooling(_GlobalPool):
    MODE = "avg"


class MinkowskiGlobalAvgPooling(_GlobalPool):
    MODE = "avg"


class MinkowskiGlobalSumPooling(_GlobalPool):
    MODE = "sum"


class MinkowskiGlobalMaxPooling(_GlobalPool):
    MODE = "max"


class MinkowskiBroadcastMultiplication(MinkowskiModuleBase):
    def forward(self, x, y):
        cm = x.coordinate_manager
        return x._wrap(oo.broadcast_mul(x.F, y.F, cm.batch_of(x.coordinate_map_key)))


class MinkowskiBroadcastAddition(MinkowskiModuleBase):
    def forward(self, x, y):
        cm = x.coordinate_manager
        return x._wrap(x.F + y.F[torch.from_numpy(cm.batch_of(x.coordinate_map_key))])


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        return x._wrap(self.linear(x.F))


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        return x._wrap(self.bn(x.F))


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    pass


class MinkowskiInstanceNorm(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x):
        raise NotImplementedError("oracle: instance norm is outside the MSENet hot path")


class MinkowskiDropout(nn.Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.module = nn.Dropout(p, inplace)

    def forward(self, x):
        return x._wrap(self.module(x.F))


class MinkowskiNonlinearityBase(nn.Module):
    MODULE = None

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.module = self.MODULE(*args, **kwargs)

    def forward(self, x):
        return x._wrap(self.module(x.F))


def _nl(name, mod):
    return type(name, (MinkowskiNonlinearityBase,), {"MODULE": mod})


MinkowskiReLU = _nl("MinkowskiReLU", nn.ReLU)
MinkowskiPReLU = _nl("MinkowskiPReLU", nn.PReLU)
MinkowskiLeakyReLU = _nl("MinkowskiLeakyReLU", nn.LeakyReLU)
MinkowskiELU = _nl("MinkowskiELU", nn.ELU)
MinkowskiCELU = _nl("MinkowskiCELU", nn.CELU)
MinkowskiSELU = _nl("MinkowskiSELU", nn.SELU)
MinkowskiSiLU = _nl("MinkowskiSiLU", nn.SiLU)
MinkowskiGELU = _nl("MinkowskiGELU", nn.GELU)
MinkowskiSigmoid = _nl("MinkowskiSigmoid", nn.Sigmoid)
MinkowskiTanh = _nl("MinkowskiTanh", nn.Tanh)
MinkowskiSoftmax = _nl("MinkowskiSoftmax", nn.Softmax)
MinkowskiSoftplus = _nl("MinkowskiSoftplus", nn.Softplus)


class MinkowskiSinusoidal(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.kernel = nn.Parameter(torch.rand(in_channel, out_channel))
        self.bias = nn.Parameter(torch.rand(1, out_channel))
        self.coef = nn.Parameter(torch.rand(1, out_channel))

    def forward(self, x):
        return x._wrap(self.coef * torch.sin(x.F.mm(self.kernel) + self.bias))


def cat(*tensors):
    first = tensors[0]
    for t in tensors[1:]:
        assert t.coordinate_map_key == first.coordinate_map_key
    return first._wrap(torch.cat([t.F for t in tensors], 1))


# ---- sub-namespaces the reference imports by name (SENet.py:5, common.py:9) -----------------
_this = sys.modules[__name__]

MinkowskiNormalization = types.ModuleType("MinkowskiEngine.MinkowskiNormalization")
for _n in ("MinkowskiBatchNorm", "MinkowskiSyncBatchNorm", "MinkowskiInstanceNorm"):
    setattr(MinkowskiNormalization, _n, getattr(_this, _n))

MinkowskiNonlinearity = types.ModuleType("MinkowskiEngine.MinkowskiNonlinearity")
for _n in ("MinkowskiReLU", "MinkowskiPReLU", "MinkowskiLeakyReLU", "MinkowskiELU", "MinkowskiCELU",
           "MinkowskiSELU", "MinkowskiSiLU", "MinkowskiGELU", "MinkowskiSigmoid", "MinkowskiTanh",
           "MinkowskiSoftmax", "MinkowskiSoftplus", "MinkowskiSinusoidal", "MinkowskiNonlinearityBase"):
    setattr(MinkowskiNonlinearity, _n, getattr(_this, _n))

utils = types.ModuleType("MinkowskiEngine.utils")
utils.kaiming_normal_ = lambda tensor, a=0, mode="fan_in", nonlinearity="leaky_relu": \
    nn.init.kaiming_normal_(tensor, a=a, mode=mode, nonlinearity=nonlinearity)


def install():
    """Register this namespace as ``MinkowskiEngine`` (tests / CPU baseline only)."""
    sys.modules["MinkowskiEngine"] = _this
    sys.modules["MinkowskiEngine.MinkowskiNormalization"] = MinkowskiNormalization
    sys.modules["MinkowskiEngine.MinkowskiNonlinearity"] = MinkowskiNonlinearity
    sys.modules["MinkowskiEngine.utils"] = utils
    return _this


def uninstall():
    for k in [k for k in sys.modules if k == "MinkowskiEngine" or k.startswith("MinkowskiEngine.")]:
        del sys.modules[k]
