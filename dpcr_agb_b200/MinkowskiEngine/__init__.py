"""Drop-in ``MinkowskiEngine`` module surface for DPCR-AGB's MSENet14 / MSENet50, B200-native.

Usage from unchanged reference code::

    import dpcr_agb_b200
    dpcr_agb_b200.install()            # registers this package as ``MinkowskiEngine``
    import MinkowskiEngine as ME       # torch_points3d/modules/MinkowskiEngine/*.py now run on libb200sparse

Everything underneath is hand-written sm_100a CUDA behind the C ABI of ``include/b200sparse.h``;
there is no CPU path (SparseTensor creation raises on CPU tensors).
"""
from .coordinate_manager import CoordinateManager, CoordinateMapKey, KernelMap  # noqa: F401
from .sparse_tensor import SparseTensor  # noqa: F401
from .modules import *  # noqa: F401,F403
from .modules import (KernelGenerator, MinkowskiNetwork, PoolingMode, RegionType, cat)  # noqa: F401
from . import MinkowskiNonlinearity, MinkowskiNormalization, utils  # noqa: F401
from . import functional as MinkowskiFunctional  # noqa: F401

__version__ = "0.5.4+b200"

# Extensions beyond the MinkowskiEngine surface, used by dpcr_agb_b200.msenet when present (the unchanged
# reference networks never touch them): MinkowskiBatchNorm.forward(x, act=1) fuses the exact GELU into the
# batch-norm apply kernel, fused_add_gelu(x, y) is the residual join act(x + y) as one kernel.
B200_FUSED_OPS = True


def fused_add_gelu(x, y, tf32_out=0):
    """``tf32_out``: 1 = also emit the TF32 operand twin for the convolutions that consume the result, 2 = the result
    is consumed by convolutions only and is written TF32-rounded (dpcr_agb_b200.msenet decides per block)."""
    assert x._same_map(y), "fused_add_gelu needs both tensors on the same coordinate map"
    Fn = MinkowskiFunctional
    if not Fn.twins_on() or x.F.shape[1] <= 4:
        tf32_out = 0
    if tf32_out == 1:
        out, out_r = Fn.AddGELUFunction.apply(x.F, y.F, x.n_dev, 1)
        Fn.attach_twin(out, out_r)
    else:
        out = Fn.AddGELUFunction.apply(x.F, y.F, x.n_dev, tf32_out)
        if tf32_out == 2:
            Fn.mark_rounded(out)
    return x._wrap(out)


def fused_se_tail(u, res, fc1, fc2, keep=None, tf32_out=0):
    """``gelu(u * (sigmoid(fc2(gelu(fc1(mean_plot(u))))) * keep[plot]) + res)`` -- squeeze-excite gate, drop-path
    scale, residual join and activation of an SE block as one op (``MinkowskiFunctional.SETailFunction``).
    ``fc1`` / ``fc2``: the block's ``MinkowskiLinear`` modules; ``keep``: [B,1] drop-path scale or None."""
    assert u._same_map(res), "fused_se_tail needs both tensors on the same coordinate map"
    Fn = MinkowskiFunctional
    cm, key = u.coordinate_manager, u.coordinate_map_key
    if not Fn.twins_on() or u.F.shape[1] <= 4:
        tf32_out = 0
    args = (u.F, res.F, fc1.linear.weight, fc1.linear.bias, fc2.linear.weight, fc2.linear.bias, keep, cm.coords(key),
            cm.num_batches, cm.inv_counts(key), u.n_dev, tf32_out)
    if tf32_out == 1:
        out, out_r = Fn.SETailFunction.apply(*args)
        Fn.attach_twin(out, out_r)
    else:
        out = Fn.SETailFunction.apply(*args)
        if tf32_out == 2:
            Fn.mark_rounded(out)
    return u._wrap(out)
