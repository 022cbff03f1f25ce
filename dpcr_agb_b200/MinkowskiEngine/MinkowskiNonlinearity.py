"""``from MinkowskiEngine import MinkowskiNonlinearity as NL`` (reference: common.py:9,32-42)."""
from .modules import (MinkowskiCELU, MinkowskiELU, MinkowskiGELU, MinkowskiHardshrink,  # noqa: F401
                      MinkowskiHardsigmoid, MinkowskiHardswish, MinkowskiHardtanh, MinkowskiLeakyReLU,
                      MinkowskiLogSigmoid, MinkowskiLogSoftmax, MinkowskiNonlinearityBase, MinkowskiPReLU,
                      MinkowskiReLU, MinkowskiReLU6, MinkowskiRReLU, MinkowskiSELU, MinkowskiSigmoid,
                      MinkowskiSiLU, MinkowskiSinusoidal, MinkowskiSoftmax, MinkowskiSoftmin,
                      MinkowskiSoftplus, MinkowskiSoftshrink, MinkowskiSoftsign, MinkowskiTanh,
                      MinkowskiTanhshrink, MinkowskiThreshold)
