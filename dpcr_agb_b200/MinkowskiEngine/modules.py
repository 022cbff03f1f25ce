"""nn.Module surface of MinkowskiEngine that DPCR-AGB's networks import.

Parameter / attribute names are the ones the reference touches -- ``.kernel`` / ``.bias`` on
convolutions (``SENet.py:80-83``), ``.bn`` on batch norm (``SENet.py:76-78``,
``core/schedulers/bn_schedulers.py:13``), ``.linear`` on linear (``SENet.py:84-87``,
``models/instance/minkowski.py:39``) -- so that released checkpoints keep loading.
"""
from __future__ import annotations

import math
from enum import Enum

import torch
import torch.nn as nn

from . import functional as Fn
from .coordinate_manager import _triple
from .sparse_tensor import SparseTensor


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class PoolingMode(Enum):
    LOCAL_SUM_POOLING = 0
    LOCAL_AVG_POOLING = 1
    LOCAL_MAX_POOLING = 2
    GLOBAL_SUM_POOLING_DEFAULT = 3
    GLOBAL_AVG_POOLING_DEFAULT = 4
    GLOBAL_MAX_POOLING_DEFAULT = 5
    GLOBAL_SUM_POOLING_KERNEL = 6
    GLOBAL_AVG_POOLING_KERNEL = 7
    GLOBAL_MAX_POOLING_KERNEL = 8
    GLOBAL_SUM_POOLING_PYTORCH_INDEX = 9
    GLOBAL_AVG_POOLING_PYTORCH_INDEX = 10
    GLOBAL_MAX_POOLING_PYTORCH_INDEX = 11


class KernelGenerator:
    """Only hyper-cube regions are generated (the reference's README configs use nothing else)."""

    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False,
                 region_type=RegionType.HYPER_CUBE, region_offsets=None, expand_coordinates=False,
                 axis_types=None, dimension=-1):
        assert dimension == 3, "D=3 only"
        self.kernel_size = _triple(kernel_size)
        self.kernel_stride = _triple(stride)
        self.kernel_dilation = _triple(dilation)
        self.region_type = region_type
        self.is_transpose = is_transpose
        self.expand_coordinates = expand_coordinates
        self.kernel_volume = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        self.dimension = dimension
        if region_type != RegionType.HYPER_CUBE:
            raise NotImplementedError("only RegionType.HYPER_CUBE kernels are built (SURVEY.md 8f)")


class MinkowskiModuleBase(nn.Module):
    pass


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


def _out_key(x: SparseTensor, stride):
    cm, in_key = x.coordinate_manager, x.coordinate_map_key
    if stride == (1, 1, 1):
        return in_key
    return cm.stride(in_key, stride)


class MinkowskiConvolution(MinkowskiModuleBase):
    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3, "dimension=3 is the case DPCR-AGB uses and the one implemented"
        if kernel_generator is not None:
            kernel_size, stride, dilation = (kernel_generator.kernel_size, kernel_generator.kernel_stride,
                                             kernel_generator.kernel_dilation)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)
        self.kernel_volume = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        self.dimension = dimension
        self.is_transpose = False
        # K=1 and stride=1: plain [N,Cin] @ [Cin,Cout] on the same coordinate map (ME's ``use_mm``)
        self.use_mm = self.kernel_volume == 1 and self.stride == (1, 1, 1)
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.empty((1, out_channels), dtype=torch.float32)) if bias else None
        self.reset_parameters()

    def reset_parameters(self, is_transpose=False):
        with torch.no_grad():
            n = (self.out_channels if is_transpose else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        # every convolution of the reference's networks feeds a batch norm: its statistics ride in the epilogue
        if self.use_mm:
            stats = Fn.new_col_stats(input.F.shape[0], self.out_channels, input.F.device) if self.training else None
            out = Fn.ConvolutionFunction.apply(input.F, self.kernel, self.bias, None, input.n_dev, stats)
            Fn.attach_col_stats(out, stats)
            return SparseTensor(out, coordinate_map_key=input.coordinate_map_key, coordinate_manager=cm)
        out_key = _out_key(input, self.stride)
        kmap = cm.kernel_map(input.coordinate_map_key, out_key, self.kernel_size, self.dilation)
        stats = Fn.new_col_stats(kmap.n_out, self.out_channels, input.F.device) if self.training else None
        out = Fn.ConvolutionFunction.apply(input.F, self.kernel, self.bias, kmap, None, stats)
        Fn.attach_col_stats(out, stats)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)

    def __repr__(self):
        return (f"{type(self).__name__}(in={self.in_channels}, out={self.out_channels}, "
                f"kernel_size={list(self.kernel_size)}, stride={list(self.stride)}, dilation={list(self.dilation)})")


class _NotOnHotPath(MinkowskiModuleBase):
    """Names that must exist for the reference package to import but that none of its networks call."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} is not used by any network of the reference")


def _fine_key(x: SparseTensor, stride):
    """Key of the EXISTING finer map a transposed op lands on (tensor stride / stride, same tag): the encoder map of a
    U-Net (``networks.py:155-245``).  Generating new coordinates (``expand_coordinates``) is not implemented."""
    cm, in_key = x.coordinate_manager, x.coordinate_map_key
    ts = in_key.tensor_stride
    if any(t % s for t, s in zip(ts, stride)):
        raise ValueError(f"tensor stride {ts} is not divisible by the transposed stride {stride}")
    from .coordinate_manager import CoordinateMapKey
    key = CoordinateMapKey(tuple(t // s for t, s in zip(ts, stride)), in_key.tag)
    if key not in cm.maps:
        raise NotImplementedError(f"transposed op onto tensor stride {list(key.tensor_stride)}: no such coordinate map "
                                  f"in this manager (generative transposed convolution is not implemented)")
    return key


class MinkowskiConvolutionTranspose(MinkowskiConvolution):
    """Non-generative transposed convolution (``networks.py:155-176``: the decoder of MinkUNet): the output lives on
    the existing finer map; ``out[f] = bias + sum_k in[c] @ W[k]`` over the pairs of the forward convolution
    fine -> coarse with the two sides swapped.  Runs on the same gather-GEMM / wgrad kernels as the convolution: its
    forward table is the transposed table of that forward map (``KernelMap.transposed``)."""

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        if expand_coordinates:
            raise NotImplementedError("expand_coordinates (generative transposed convolution) is not implemented")
        super().__init__(in_channels, out_channels, kernel_size, stride, dilation, bias, kernel_generator, False,
                         convolution_mode, dimension)
        self.is_transpose = True
        if self.use_mm:          # K=1, stride=1: the transposed op is the same matmul; keep ME's [K^3, Cin, Cout] shape
            self.use_mm = False
            self.kernel = nn.Parameter(torch.empty((self.kernel_volume, in_channels, out_channels)))
        self.reset_parameters(True)

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        out_key = _fine_key(input, self.stride)
        fwd_map = cm.kernel_map(out_key, input.coordinate_map_key, self.kernel_size, self.dilation)
        out = Fn.ConvolutionFunction.apply(input.F, self.kernel, self.bias, fwd_map.transposed())
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiSumPooling(MinkowskiModuleBase):
    """Local sum pooling over the kernel region; ``AVERAGE`` (MinkowskiAvgPooling, ``networks.py:29``) divides by the
    number of inputs under the kernel, as MinkowskiEngine does."""
    AVERAGE = False

    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        assert dimension == 3
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        out_key = _out_key(input, self.stride)
        kmap = cm.kernel_map(input.coordinate_map_key, out_key, self.kernel_size, self.dilation)
        out = Fn.LocalPoolFunction.apply(input.F, kmap, self.AVERAGE)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)

    def __repr__(self):
        return f"{type(self).__name__}(kernel_size={list(self.kernel_size)}, stride={list(self.stride)})"


class MinkowskiAvgPooling(MinkowskiSumPooling):
    AVERAGE = True


class MinkowskiPoolingTranspose(MinkowskiSumPooling):
    """Unpooling onto the existing finer map: ``out[f] = sum_k in[c]`` over the transposed pairs of the forward
    pooling fine -> coarse."""

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        out_key = _fine_key(input, self.stride)
        fwd_map = cm.kernel_map(out_key, input.coordinate_map_key, self.kernel_size, self.dilation)
        out = Fn.LocalPoolFunction.apply(input.F, fwd_map.transposed(), False)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiAvgUnpooling(_NotOnHotPath):
    pass


class MinkowskiMaxPooling(MinkowskiModuleBase):
    def __init__(self, kernel_size, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        assert dimension == 3
        self.kernel_size, self.stride, self.dilation = _triple(kernel_size), _triple(stride), _triple(dilation)

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        out_key = _out_key(input, self.stride)
        kmap = cm.kernel_map(input.coordinate_map_key, out_key, self.kernel_size, self.dilation)
        if Fn.twins_on() and input.F.shape[1] > 4:      # the pooled rows feed a convolution: TF32 operand alongside
            out, out_r = Fn.MaxPoolFunction.apply(input.F, kmap, True)
            Fn.attach_twin(out, out_r)
        else:
            out = Fn.MaxPoolFunction.apply(input.F, kmap)
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)

    def __repr__(self):
        return f"{type(self).__name__}(kernel_size={list(self.kernel_size)}, stride={list(self.stride)})"


class _GlobalPoolBase(MinkowskiModuleBase):
    AVERAGE = False

    def __init__(self, mode=None):
        super().__init__()
        self.mode = mode

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        key = input.coordinate_map_key
        scale = cm.inv_counts(key) if self.AVERAGE else None
        out = Fn.GlobalPoolFunction.apply(input.F, cm.coords(key), cm.num_batches, scale, cm.n_dev(key))
        return SparseTensor(out, coordinate_map_key=cm.origin(), coordinate_manager=cm)

    def __repr__(self):
        return type(self).__name__ + "()"


class MinkowskiGlobalSumPooling(_GlobalPoolBase):
    AVERAGE = False


class MinkowskiGlobalAvgPooling(_GlobalPoolBase):
    AVERAGE = True


class MinkowskiGlobalPooling(_GlobalPoolBase):
    """ME's default global pooling mode is average (the SE "squeeze", ``senet_block.py:43``)."""
    AVERAGE = True


class MinkowskiGlobalMaxPooling(MinkowskiModuleBase):
    """Per-plot maximum (``networks.py:39``, ``PointNet.py:28``): ``b2s_segment_max`` over the batch-sorted rows."""

    def __init__(self, mode=None):
        super().__init__()

    def forward(self, input: SparseTensor, coordinates=None):
        cm = input.coordinate_manager
        key = input.coordinate_map_key
        out = Fn.GlobalMaxPoolFunction.apply(input.F, cm.coords(key), cm.num_batches, cm.n_dev(key))
        return SparseTensor(out, coordinate_map_key=cm.origin(), coordinate_manager=cm)


class MinkowskiBroadcastMultiplication(MinkowskiModuleBase):
    def forward(self, input: SparseTensor, input_glob: SparseTensor):
        cm = input.coordinate_manager
        out = Fn.BroadcastMulFunction.apply(input.F, input_glob.F, cm.coords(input.coordinate_map_key), cm.num_batches,
                                            input.n_dev)
        return input._wrap(out)

    def __repr__(self):
        return type(self).__name__ + "()"


class MinkowskiBroadcast(MinkowskiModuleBase):
    def forward(self, input: SparseTensor, input_glob: SparseTensor):
        cm = input.coordinate_manager
        key = input.coordinate_map_key
        return input._wrap(Fn.BroadcastFunction.apply(input_glob.F, cm.coords(key), input.F.shape[0], cm.n_dev(key)))


class MinkowskiBroadcastAddition(MinkowskiBroadcast):
    def forward(self, input: SparseTensor, input_glob: SparseTensor):
        return input._wrap(input.F + super().forward(input, input_glob).F)


class MinkowskiLinear(nn.Module):
    """nn.Linear on ``.F`` -- the SE MLP ([B,C] x [C,C/16], negligible FLOPs) and the head."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input):
        return input._wrap(self.linear(input.F))

    def __repr__(self):
        return (f"{type(self).__name__}(in_features={self.linear.in_features}, "
                f"out_features={self.linear.out_features}, bias={self.linear.bias is not None})")


# ``with deferred_bn_counters():`` (dpcr_agb_b200.train) -- the ``num_batches_tracked += 1`` of every batch-norm layer
# becomes ONE multi-tensor add after the forward pass instead of one tiny launch per layer.
DEFERRED_BN_COUNTERS = None


class deferred_bn_counters:
    def __enter__(self):
        global DEFERRED_BN_COUNTERS
        self.old, DEFERRED_BN_COUNTERS = DEFERRED_BN_COUNTERS, []

    def __exit__(self, *exc):
        global DEFERRED_BN_COUNTERS
        pending, DEFERRED_BN_COUNTERS = DEFERRED_BN_COUNTERS, self.old
        if pending and exc[0] is None:
            torch._foreach_add_(pending, 1)


class MinkowskiBatchNorm(nn.Module):
    """nn.BatchNorm1d semantics over all rows of the batch, computed by the b2s_bn_* kernels.

    ``self.bn`` is a real ``nn.BatchNorm1d`` so that state-dict keys (``bn.weight`` ...), ``momentum``
    updates by the BN scheduler (``bn_schedulers.py:28-30``) and ``init_weights`` keep working."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, input: SparseTensor, act: int = 0, tf32_only: bool = False):
        """``act=1`` fuses the exact GELU; ``tf32_only`` (product-path extension, used by dpcr_agb_b200.msenet when
        the only consumer is a convolution) writes the result already rounded to the TF32 operand form."""
        bn = self.bn
        tf32_only = bool(tf32_only and Fn.twins_on() and input.F.shape[1] > 4)
        use_batch_stats = bn.training or bn.running_mean is None
        momentum = bn.momentum
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            if DEFERRED_BN_COUNTERS is not None and momentum is not None:
                DEFERRED_BN_COUNTERS.append(bn.num_batches_tracked)   # one fused add for all layers after the forward
            else:
                bn.num_batches_tracked.add_(1)
            if momentum is None:
                momentum = 1.0 / float(bn.num_batches_tracked)
        update = bn.training and bn.track_running_stats
        out = Fn.BatchNormFunction.apply(input.F, bn.weight, bn.bias,
                                         bn.running_mean if (update or not use_batch_stats) else None,
                                         bn.running_var if (update or not use_batch_stats) else None,
                                         use_batch_stats, 0.0 if momentum is None else momentum, bn.eps, act,
                                         input.n_dev, tf32_only,
                                         Fn.col_stats_of(input.F, input.F.shape[1]) if use_batch_stats else None)
        if tf32_only:
            Fn.mark_rounded(out)
        return input._wrap(out)

    def __repr__(self):
        b = self.bn
        return (f"{type(self).__name__}({b.num_features}, eps={b.eps}, momentum={b.momentum}, "
                f"affine={b.affine}, track_running_stats={b.track_running_stats})")


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    """Per-replica statistics, as in the reference's DataParallel intent."""


class MinkowskiInstanceNorm(nn.Module):
    """Exists so that the isinstance tuple at ``bn_schedulers.py:9-15`` builds; ``norm_type="in"`` is not a
    README configuration."""

    def __init__(self, num_features):
        super().__init__()
        self.num_features = num_features
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, input):
        raise NotImplementedError("MinkowskiInstanceNorm is outside the MSENet hot path")


class MinkowskiStableInstanceNorm(MinkowskiInstanceNorm):
    pass


class MinkowskiDropout(nn.Module):
    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.module = nn.Dropout(p, inplace)

    def forward(self, input):
        return input._wrap(self.module(input.F))


class MinkowskiNonlinearityBase(nn.Module):
    MODULE = None

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.module = self.MODULE(*args, **kwargs)

    def forward(self, input):
        return input._wrap(self.module(input.F))

    def __repr__(self):
        return type(self).__name__ + "()"


def _nonlinearity(name, torch_module):
    return type(name, (MinkowskiNonlinearityBase,), {"MODULE": torch_module, "__module__": __name__})


MinkowskiELU = _nonlinearity("MinkowskiELU", nn.ELU)
MinkowskiHardshrink = _nonlinearity("MinkowskiHardshrink", nn.Hardshrink)
MinkowskiHardsigmoid = _nonlinearity("MinkowskiHardsigmoid", nn.Hardsigmoid)
MinkowskiHardtanh = _nonlinearity("MinkowskiHardtanh", nn.Hardtanh)
MinkowskiHardswish = _nonlinearity("MinkowskiHardswish", nn.Hardswish)
MinkowskiLeakyReLU = _nonlinearity("MinkowskiLeakyReLU", nn.LeakyReLU)
MinkowskiLogSigmoid = _nonlinearity("MinkowskiLogSigmoid", nn.LogSigmoid)
MinkowskiPReLU = _nonlinearity("MinkowskiPReLU", nn.PReLU)
MinkowskiReLU = _nonlinearity("MinkowskiReLU", nn.ReLU)
MinkowskiReLU6 = _nonlinearity("MinkowskiReLU6", nn.ReLU6)
MinkowskiRReLU = _nonlinearity("MinkowskiRReLU", nn.RReLU)
MinkowskiSELU = _nonlinearity("MinkowskiSELU", nn.SELU)
MinkowskiCELU = _nonlinearity("MinkowskiCELU", nn.CELU)
MinkowskiSigmoid = _nonlinearity("MinkowskiSigmoid", nn.Sigmoid)
MinkowskiSiLU = _nonlinearity("MinkowskiSiLU", nn.SiLU)
MinkowskiSoftplus = _nonlinearity("MinkowskiSoftplus", nn.Softplus)
MinkowskiSoftshrink = _nonlinearity("MinkowskiSoftshrink", nn.Softshrink)
MinkowskiSoftsign = _nonlinearity("MinkowskiSoftsign", nn.Softsign)
MinkowskiTanh = _nonlinearity("MinkowskiTanh", nn.Tanh)
MinkowskiTanhshrink = _nonlinearity("MinkowskiTanhshrink", nn.Tanhshrink)
MinkowskiThreshold = _nonlinearity("MinkowskiThreshold", nn.Threshold)
MinkowskiSoftmin = _nonlinearity("MinkowskiSoftmin", nn.Softmin)
MinkowskiSoftmax = _nonlinearity("MinkowskiSoftmax", nn.Softmax)
MinkowskiLogSoftmax = _nonlinearity("MinkowskiLogSoftmax", nn.LogSoftmax)


class MinkowskiGELU(MinkowskiNonlinearityBase):
    """Exact-erf GELU through the b2s_gelu kernels (the activation of every README config,
    ``conf/models/instance/minkowski_baseline.yaml:76``)."""
    MODULE = nn.GELU

    def forward(self, input):
        if getattr(self.module, "approximate", "none") != "none":
            return input._wrap(self.module(input.F))
        return input._wrap(Fn.GELUFunction.apply(input.F, input.n_dev))


class MinkowskiSinusoidal(MinkowskiModuleBase):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.in_channel, self.out_channel = in_channel, out_channel
        self.kernel = nn.Parameter(torch.rand(in_channel, out_channel))
        self.bias = nn.Parameter(torch.rand(1, out_channel))
        self.coef = nn.Parameter(torch.rand(1, out_channel))

    def forward(self, input):
        return input._wrap(self.coef * torch.sin(input.F.mm(self.kernel) + self.bias))


def cat(*sparse_tensors):
    first = sparse_tensors[0]
    for t in sparse_tensors[1:]:
        if not first._same_map(t):
            raise NotImplementedError("ME.cat across coordinate maps is outside the MSENet hot path")
    return first._wrap(torch.cat([t.F for t in sparse_tensors], dim=1))
