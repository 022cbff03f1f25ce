"""Floating-point half of the oracle: the arithmetic MinkowskiEngine performs for MSENet14/50.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED -- restates the
MinkowskiEngine (~v0.5.4, CPU build ``env_cpu.yml``) algorithm "gather -> GEMM -> scatter per
kernel offset"; gradients are whatever torch autograd derives from that restatement.

Every function takes the neighbour table of ``oracle.coords.kernel_map_table`` and plain torch CPU
tensors (fp32 by default; pass fp64 tensors for a higher-precision yardstick).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _as_long(a):
    if isinstance(a, np.ndarray):
        return torch.from_numpy(a.astype(np.int64))
    return a.long()


def conv(x: torch.Tensor, weight: torch.Tensor, nbr, bias: torch.Tensor | None = None) -> torch.Tensor:
    """MinkowskiConvolution forward: ``out[o] = bias + sum_k sum_{(i->o) in M_k} x[i] @ W[k]``.

    Call sites in the reference: ``modules/MinkowskiEngine/SENet.py:49-52,94-97``,
    ``resnet_block.py:48-54,95-107``, ``common.py:219-221``.  ``weight`` is ``[K^3, Cin, Cout]``
    (or ``[Cin, Cout]`` for the K=1, stride=1 ``use_mm`` case, where ``nbr`` is ignored)."""
    if weight.dim() == 2:
        out = x @ weight
    else:
        nbr = _as_long(nbr)
        n_out = nbr.shape[1]
        out = x.new_zeros((n_out, weight.shape[2]))
        for k in range(nbr.shape[0]):
            o = torch.nonzero(nbr[k] >= 0).squeeze(1)
            if o.numel() == 0:
                continue
            out = out.index_add(0, o, x[nbr[k, o]] @ weight[k])
    if bias is not None:
        out = out + bias
    return out


def max_pool(x: torch.Tensor, nbr) -> torch.Tensor:
    """MinkowskiMaxPooling forward (``SENet.py:53``): max over the rows the kernel map pairs with an
    out row; ties resolve to the lowest in-row; backward routes the gradient to that row only."""
    nbr = _as_long(nbr)
    k3, n_out = nbr.shape
    c = x.shape[1]
    valid = nbr >= 0
    g = x[nbr.clamp(min=0)]                                    # [K3, N_out, C]
    g = torch.where(valid[:, :, None], g, torch.full_like(g, -float("inf")))
    m = g.max(0).values                                        # [N_out, C]
    big = torch.iinfo(torch.int64).max
    cand = torch.where((g == m[None]) & valid[:, :, None], nbr[:, :, None].expand(k3, n_out, c),
                       torch.full((1, 1, 1), big, dtype=torch.int64))
    arg = cand.min(0).values                                   # lowest in-row among the maxima
    has = valid.any(0)
    arg = torch.where(has[:, None], arg, torch.zeros_like(arg))
    out = torch.gather(x, 0, arg)
    return torch.where(has[:, None], out, torch.zeros_like(out))


def global_pool(x: torch.Tensor, batch, num_batches: int, mode: str = "avg") -> torch.Tensor:
    """MinkowskiGlobal{Sum,Avg,Max}Pooling / MinkowskiGlobalPooling (``senet_block.py:43``,
    ``common.py:44-48``): per-plot reduction ``[N,C] -> [B,C]``, rows ordered by batch id."""
    b = _as_long(batch)
    if mode == "max":
        out = x.new_full((num_batches, x.shape[1]), -float("inf"))
        return out.scatter_reduce(0, b[:, None].expand_as(x), x, reduce="amax", include_self=True)
    out = x.new_zeros((num_batches, x.shape[1])).index_add(0, b, x)
    if mode == "avg":
        cnt = torch.bincount(b, minlength=num_batches).clamp(min=1).to(x.dtype)
        out = out / cnt[:, None]
    return out


def broadcast_mul(x: torch.Tensor, y: torch.Tensor, batch) -> torch.Tensor:
    """MinkowskiBroadcastMultiplication (``senet_block.py:44,50``): ``out[i] = x[i] * y[batch(i)]``."""
    return x * y[_as_long(batch)]


def batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps=1e-5):
    """MinkowskiBatchNorm == nn.BatchNorm1d over all rows of the batch (``SENet.py:35``)."""
    return F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)


def gelu(x):
    """NL.MinkowskiGELU == exact-erf nn.GELU() on ``.F`` (``common.py:41``)."""
    return F.gelu(x)
