"""MSENet14 / MSENet50 (and the sibling depths) written against a MinkowskiEngine-shaped namespace.

This restates the network *definitions* of the reference --
``torch_points3d/modules/MinkowskiEngine/SENet.py:14-118,151-188`` (ResNetBase / SENet14 / SENet50),
``senet_block.py:33-147`` (SELayer, SEBasicBlock, SEBottleneck), ``resnet_block.py:31-133``,
``common.py:215-226,344-366`` (ConvNormActivation, MinkowskiDropPath) and the ``SeparateLinear`` head of
``models/instance/minkowski.py:15-26,39-46`` -- because ``/root/reference`` does not exist where the GPU
tests and the benchmark run.  Module attribute names are kept identical, so ``state_dict()`` keys match
the reference's and its checkpoints load (tests/test_api_surface.py checks key-by-key equality against
the unchanged reference classes when the reference tree is present).

``ME`` is passed in: ``dpcr_agb_b200.MinkowskiEngine`` for the CUDA product path, ``oracle.me_cpu`` for
the CPU checker.  The unchanged reference classes run over the same namespaces via ``install()``.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn

ARCH = {
    #  name        (bottleneck, layers,          init_dim, planes)
    "SENet14": (False, (1, 1, 1, 1), 64, (64, 128, 256, 512)),
    "SENet18": (False, (2, 2, 2, 2), 64, (64, 128, 256, 512)),
    "SENet34": (False, (3, 4, 6, 3), 64, (64, 128, 256, 512)),
    "SENet50": (True, (3, 4, 6, 3), 64, (64, 128, 256, 512)),
    "SENet101": (True, (3, 4, 23, 3), 64, (64, 128, 256, 512)),
}
ACTIVATION_NAMES = {"relu": "MinkowskiReLU", "gelu": "MinkowskiGELU", "silu": "MinkowskiSiLU",
                    "swish": "MinkowskiSiLU", "sigmoid": "MinkowskiSigmoid", "tanh": "MinkowskiTanh"}
POOL_NAMES = {"sum": "MinkowskiGlobalSumPooling", "mean": "MinkowskiGlobalAvgPooling",
              "max": "MinkowskiGlobalMaxPooling"}


def _rewrap(ME, x, feats):
    return ME.SparseTensor(feats, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


class DropPath(nn.Module):
    """Per-plot stochastic depth, same draws as ``common.py:353-366``: one ``random.uniform(0, 1)`` per
    plot in batch order, kept plots scaled by 1/keep.  The mask is a [B,1] tensor applied with the
    broadcast-multiply op instead of a host-built per-row mask (no device->host sync)."""

    def __init__(self, ME, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep
        self._ME = [ME]                      # list: keep the namespace out of nn.Module registration
        self.mul = ME.MinkowskiBroadcastMultiplication()
        # captured-graph mode (dpcr_agb_b200.graph_step): a persistent [B,1] device tensor that the host refills
        # with draw() before every replay, instead of a fresh tensor per call
        self.static_mask = None

    def draw(self, nb):
        """The per-plot keep/scale values of one forward call (python ``random``, the reference's RNG)."""
        keep = 1.0 - self.drop_prob
        scale = 1.0 / keep if (keep > 0.0 and self.scale_by_keep) else 1.0
        return [scale if random.uniform(0, 1) > self.drop_prob else 0.0 for _ in range(nb)]

    def mask(self, x):
        """[B,1] keep/scale tensor of this forward call on the device, or None in eval mode."""
        if not self.training:
            return None
        cm = x.coordinate_manager
        nb = cm.num_batches if isinstance(cm.num_batches, int) else cm.num_batches()
        if self.static_mask is not None:
            return self.static_mask
        return torch.tensor(self.draw(nb), dtype=x.F.dtype).view(nb, 1).to(x.F.device, non_blocking=True)

    def forward(self, x):
        mask = self.mask(x)
        if mask is None:
            return x
        ME = self._ME[0]
        cm = x.coordinate_manager
        glob = ME.SparseTensor(mask, coordinate_map_key=cm.origin(x.coordinate_map_key), coordinate_manager=cm)
        return self.mul(x, glob)


def _is_gelu(ME, act):
    return isinstance(act, ME.MinkowskiNonlinearity.MinkowskiGELU)


# Stream for the downsample branch of strided residual blocks (see SEResidualBlock._residual_branch); None = off.
# dpcr_agb_b200.train.Trainer sets it for the step it drives (B2S_BRANCH_STREAM=1).
BRANCH_STREAM = None


class ConvNormAct(nn.Module):
    def __init__(self, ME, cin, cout, kernel_size, stride, norm_layer, act, bias, fuse=False):
        super().__init__()
        self.conv = ME.MinkowskiConvolution(cin, cout, kernel_size=kernel_size, stride=stride, dimension=3, bias=bias)
        self.norm = norm_layer(cout)
        self.act = nn.Identity() if act is None else act
        self._fuse = fuse and act is not None and _is_gelu(ME, act)

    def forward(self, x):
        if self._fuse:                                   # batch norm + exact GELU in one kernel (same arithmetic)
            return self.norm(self.conv(x), act=1)
        return self.act(self.norm(self.conv(x)))


class SqueezeExcite(nn.Module):
    def __init__(self, ME, channels, act, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(ME.MinkowskiLinear(channels, channels // reduction), act,
                                ME.MinkowskiLinear(channels // reduction, channels), ME.MinkowskiSigmoid())
        self.pooling = ME.MinkowskiGlobalPooling()
        self.broadcast_mul = ME.MinkowskiBroadcastMultiplication()

    def forward(self, x):
        return self.broadcast_mul(x, self.fc(self.pooling(x)))


class SEResidualBlock(nn.Module):
    """SEBasicBlock (expansion 1: k3 - k3) or SEBottleneck (expansion 4: k1 - k3 - k1), + SE + residual."""

    def __init__(self, ME, inplanes, planes, act, norm_layer, bottleneck, stride=1, downsample=None,
                 drop_path=0.0, bias=True, fuse=False):
        super().__init__()
        self._fuse = fuse and _is_gelu(ME, act)
        self._add_act = [ME.fused_add_gelu] if self._fuse else None
        # squeeze-excite gate + drop path + residual add + GELU as one op where the namespace has it and the channel
        # count suits its kernels
        self._se_tail = [ME.fused_se_tail] if (self._fuse and hasattr(ME, "fused_se_tail")
                                               and (planes * (4 if bottleneck else 1)) % 64 == 0) else None
        # who consumes this block's output (set by MSENet): 0 = something that needs the plain values, 1 = convolutions
        # and an identity residual (plain result + TF32 operand twin), 2 = convolutions only (written TF32-rounded)
        self.out_tf32 = 0
        self.expansion = 4 if bottleneck else 1
        if bottleneck:
            spec = [(inplanes, planes, 1, 1), (planes, planes, 3, stride), (planes, planes * 4, 1, 1)]
        else:
            spec = [(inplanes, planes, 3, stride), (planes, planes, 3, 1)]
        self.num_convs = len(spec)
        for i, (ci, co, k, s) in enumerate(spec, 1):
            setattr(self, f"conv{i}", ME.MinkowskiConvolution(ci, co, kernel_size=k, stride=s, dilation=1,
                                                              dimension=3, bias=bias))
            setattr(self, f"norm{i}", norm_layer(co))
        self.relu = act
        self.downsample = downsample if downsample is not None else nn.Identity()
        self.drop_path = DropPath(ME, drop_path) if drop_path > 0.0 else nn.Identity()
        self.se = SqueezeExcite(ME, planes * self.expansion, act)

    def _residual_branch(self, x):
        """The downsample branch (k1 convolution + norm of the first block of a stage) is independent of the main
        branch until the residual join: on a GPU it runs on ``BRANCH_STREAM`` beside conv1 .. norm2 (forward, and --
        autograd replays every op on its forward stream -- backward).  Returns (residual, stream to join or None)."""
        side = BRANCH_STREAM
        if side is None or isinstance(self.downsample, nn.Identity) or not x.F.is_cuda:
            return self.downsample(x), None
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        x.F.record_stream(side)
        twin = getattr(x.F, "_b2s_tf32", None)
        if twin is not None:
            twin[0].record_stream(side)
        with torch.cuda.stream(side):
            res = self.downsample(x)
        return res, side

    @staticmethod
    def _join_branch(res, side):
        if side is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(side)
            res.F.record_stream(cur)
        return res

    def forward(self, x):
        res, side = self._residual_branch(x) if self._fuse else (None, None)
        out = x
        for i in range(1, self.num_convs + 1):
            conv, norm = getattr(self, f"conv{i}"), getattr(self, f"norm{i}")
            if i < self.num_convs:
                # fused: batch norm + GELU in one kernel, written as the TF32 operand of the next convolution
                out = norm(conv(out), act=1, tf32_only=True) if self._fuse else self.relu(norm(conv(out)))
            else:
                out = norm(conv(out))
        if self._se_tail is not None:
            keep = self.drop_path.mask(out) if isinstance(self.drop_path, DropPath) else None
            return self._se_tail[0](out, self._join_branch(res, side), self.se.fc[0], self.se.fc[2], keep, self.out_tf32)
        out = self.se(out)
        if self._fuse:
            return self._add_act[0](self.drop_path(out), self._join_branch(res, side), self.out_tf32)
        out = self.drop_path(out) + self.downsample(x)
        return self.relu(out)


class SeparateLinear(nn.Module):
    """One ``Linear(C, 1)`` per regression target (``minkowski.py:15-26``)."""

    def __init__(self, in_channel, num_reg_classes):
        super().__init__()
        self.linears = nn.ModuleList([nn.Linear(in_channel, 1, bias=True) for _ in range(num_reg_classes)])

    def forward(self, x):
        return torch.cat([lin(x.F) for lin in self.linears], 1)


class MSENet(nn.Module):
    def __init__(self, ME, name="SENet14", in_channels=3, out_channels=2, activation="gelu", first_stride=1,
                 dropout=0.0, drop_path=0.0, bn_momentum=0.1, global_pool="sum", bias=True, separate_head=True,
                 fuse=True):
        super().__init__()
        fuse = bool(fuse and getattr(ME, "B200_FUSED_OPS", False))   # product-only kernels, same arithmetic
        bottleneck, layers, init_dim, planes_list = ARCH[name]
        self.name = name
        self.act_fn = getattr(ME.MinkowskiNonlinearity, ACTIVATION_NAMES[activation])()
        norm_layer = lambda c: ME.MinkowskiNormalization.MinkowskiBatchNorm(c, momentum=bn_momentum)  # noqa: E731
        self.inplanes = init_dim
        stages = [nn.Sequential(
            ConvNormAct(ME, in_channels, init_dim, 7, first_stride, norm_layer, self.act_fn, bias, fuse=fuse),
            ME.MinkowskiMaxPooling(kernel_size=3, stride=2, dimension=3))]
        expansion = 4 if bottleneck else 1
        for planes, count, stride in zip(planes_list, layers, (1, 2, 2, 2)):
            blocks = []
            for j in range(count):
                s = stride if j == 0 else 1
                down = None
                if j == 0 and (s != 1 or self.inplanes != planes * expansion):
                    down = nn.Sequential(
                        ME.MinkowskiConvolution(self.inplanes, planes * expansion, kernel_size=1, stride=s,
                                                dimension=3, dilation=1, bias=bias),
                        norm_layer(planes * expansion))
                blocks.append(SEResidualBlock(ME, self.inplanes, planes, self.act_fn, norm_layer, bottleneck,
                                              stride=s, downsample=down, drop_path=drop_path, bias=bias, fuse=fuse))
                self.inplanes = planes * expansion
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.ModuleList(stages)
        res_blocks = [b for st in stages[1:] for b in st]
        for blk, nxt in zip(res_blocks, res_blocks[1:]):
            blk.out_tf32 = 1 if isinstance(nxt.downsample, nn.Identity) else 2
        self.glob_avg = getattr(ME, POOL_NAMES[global_pool])()
        if dropout > 0:
            self.glob_avg = nn.Sequential(self.glob_avg, ME.MinkowskiDropout(dropout))
        if separate_head:
            self.final = SeparateLinear(self.inplanes, out_channels)
        else:
            self.final = ME.MinkowskiLinear(self.inplanes, out_channels, bias=True)
        self._init_weights(ME)

    def _init_weights(self, ME):
        """``ResNetBase.init_weights`` (SENet.py:74-87) + head init of ``minkowski.py:43-46``."""
        for m in self.modules():
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)
            elif isinstance(m, ME.MinkowskiConvolution):
                nn.init.trunc_normal_(m.kernel, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, ME.MinkowskiLinear):
                nn.init.trunc_normal_(m.linear.weight, std=0.02)
                if m.linear.bias is not None:
                    nn.init.constant_(m.linear.bias, 0)
            elif isinstance(m, SeparateLinear):
                for lin in m.linears:
                    nn.init.trunc_normal_(lin.weight, std=0.02)
                    nn.init.constant_(lin.bias, 0)

    def forward(self, x):
        for stage in self.blocks:
            x = stage(x)
        return self.final(self.glob_avg(x))


def build(ME, name="SENet14", **kw):
    """README configuration of the reference (``conf/models/instance/minkowski_baseline.yaml:71-80``)."""
    cfg = dict(in_channels=3, out_channels=2, activation="gelu", first_stride=1, dropout=0.0, drop_path=0.01,
               global_pool="sum", bias=True)
    cfg.update(kw)
    return MSENet(ME, name, **cfg)
