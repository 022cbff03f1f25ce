"""dpcr_agb_b200 -- B200-native sparse-convolution hot path for StefOe/DPCR-AGB (MSENet14 / MSENet50).

Sub-modules
-----------
lib               ctypes binding of libb200sparse.so (C ABI: include/b200sparse.h)
MinkowskiEngine   the ``ME.*`` module surface the reference's networks import
quantize          GPU GridSampling3D(mode="last") (voxel quantisation)
msenet            restatement of the reference's SENet14 / SENet50 definitions over that surface
plots             synthetic Danish-NFI-shaped LiDAR plots (benchmark input)
train             training step (loss, fused AdaBelief, DDP wiring)
"""
import sys

__version__ = "0.1.0"


def install():
    """Register ``dpcr_agb_b200.MinkowskiEngine`` as the top-level module ``MinkowskiEngine`` so that the
    reference's ``import MinkowskiEngine as ME`` (SENet.py:3, common.py:6, minkowski.py:3) resolves to it."""
    from . import MinkowskiEngine as _me
    sys.modules["MinkowskiEngine"] = _me
    sys.modules["MinkowskiEngine.MinkowskiNormalization"] = _me.MinkowskiNormalization
    sys.modules["MinkowskiEngine.MinkowskiNonlinearity"] = _me.MinkowskiNonlinearity
    sys.modules["MinkowskiEngine.utils"] = _me.utils
    return _me
