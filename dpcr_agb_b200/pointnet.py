"""``MinkowskiPointNet`` of the reference (``torch_points3d/modules/MinkowskiEngine/PointNet.py:9-49``) written
against a MinkowskiEngine-shaped namespace, like :mod:`dpcr_agb_b200.msenet` (the reference tree does not exist where
the GPU tests run).  Attribute names -- ``blocks``, ``global_pool``, ``mlp``, ``dp1``, ``final`` -- and the layer order
inside the two ``nn.Sequential`` stacks are the reference's, so ``state_dict()`` keys match (pinned against the
unchanged class over the oracle in ``tests/test_api_surface.py``).  Every layer is a per-row linear / batch norm /
activation on ``.F`` plus one per-plot pooling: on the B200 path those are cuBLAS-free ``nn.Linear`` on the feature
matrix, the ``b2s_bn_*`` kernels and ``b2s_segment_{sum,max}``.
"""
from __future__ import annotations

import torch.nn as nn

from .msenet import ACTIVATION_NAMES, POOL_NAMES


class MinkowskiPointNet(nn.Module):
    def __init__(self, ME, in_channels, out_channels, activation="relu", global_pool="max", embedding_channel=1024,
                 D=3, dropout=0.0, bn_momentum=0.1, **kwargs):
        super().__init__()
        self.act_fn = getattr(ME.MinkowskiNonlinearity, ACTIVATION_NAMES[activation])()
        lin, bn = ME.MinkowskiLinear, ME.MinkowskiBatchNorm
        self.blocks = nn.Sequential(
            lin(D + in_channels, 64, bias=False), bn(64, momentum=bn_momentum), self.act_fn,
            lin(64, 128, bias=False), bn(128, momentum=bn_momentum), self.act_fn,
            lin(128, embedding_channel, bias=False), bn(embedding_channel, momentum=bn_momentum), self.act_fn)
        self.global_pool = getattr(ME, POOL_NAMES[global_pool])()
        self.mlp = nn.Sequential(
            lin(embedding_channel, 512, bias=False), bn(512, momentum=bn_momentum), self.act_fn,
            lin(512, 256, bias=False), bn(256, momentum=bn_momentum), self.act_fn)
        self.dp1 = ME.MinkowskiDropout(dropout)
        self.final = lin(256, out_channels, bias=True)

    def forward(self, x):
        x = self.blocks(x)
        x = self.global_pool(x)
        x = self.mlp(x)
        x = self.dp1(x)
        return self.final(x)
