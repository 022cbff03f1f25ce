"""``from MinkowskiEngine import MinkowskiNormalization as N`` (reference: SENet.py:5)."""
from .modules import (MinkowskiBatchNorm, MinkowskiInstanceNorm, MinkowskiStableInstanceNorm,  # noqa: F401
                      MinkowskiSyncBatchNorm)
