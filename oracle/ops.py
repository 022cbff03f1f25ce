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


# Precision model of the convolution.  "fp32" is the reference arithmetic (MinkowskiEngine computes these
# products in fp32).  "tf32" restates what the B200 kernels compute: both operands of every product rounded to
# TF32 (10-bit mantissa, round-to-nearest, ties away from zero == PTX cvt.rna.tf32.f32), products and sums in
# fp32 -- in the forward (x, W), in dgrad (grad_out, W) and in wgrad (x, grad_out).  The parity tests hold the
# CUDA path to 1e-3 against "fp32" per op (the north_star tolerance) and to a much tighter bound against "tf32".
# "bf16x2" restates the kernels' default ("precise") operand mode: every operand v travels as the bf16 pair
# h = bf16(v), l = bf16(v - h) (round-to-nearest-even, 16-17 significant bits) and a product a*b is evaluated as
# ah*bh + al*bh + ah*bl with fp32 accumulation -- in the forward (x, W), in dgrad (grad_out, W) and in wgrad
# (x, grad_out).  Where c_in <= 4 (the k7 stem) the weights are split three ways, W = H + M + L, and the forward is
# (xh + xl)*(H + M) + xh*L; its wgrad keeps all four terms (xh + xl) * (gh + gl).
CONV_PRECISION = "fp32"


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32 on every element (fp32 in, fp32 out with the low 13 mantissa bits cleared)."""
    if t.dtype != torch.float32:
        return t
    bits = t.detach().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def split_bf16(t: torch.Tensor):
    """(h, l) with h = bf16(t), l = bf16(t - h), both returned as fp32 tensors."""
    t = t.detach()
    h = t.to(torch.bfloat16).to(torch.float32)
    l = (t - h).to(torch.bfloat16).to(torch.float32)
    return h, l


def _conv_fp32(x, weight, nbr):
    if weight.dim() == 2:
        return x @ weight
    n_out = nbr.shape[1]
    out = x.new_zeros((n_out, weight.shape[2]))
    for k in range(nbr.shape[0]):
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)
        if o.numel() == 0:
            continue
        out = out.index_add(0, o, x[nbr[k, o]] @ weight[k])
    return out


class _ConvTF32(torch.autograd.Function):
    """The same sum with operands rounded to TF32 in all three passes (see CONV_PRECISION)."""

    @staticmethod
    def forward(ctx, x, weight, nbr):
        ctx.save_for_backward(x, weight)
        ctx.nbr = nbr
        return _conv_fp32(round_tf32(x), round_tf32(weight), nbr)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gyr = round_tf32(gy)
        gx = gw = None
        with torch.enable_grad():
            if ctx.needs_input_grad[0]:
                xv = torch.zeros_like(x).requires_grad_()
                (gx,) = torch.autograd.grad(_conv_fp32(xv, round_tf32(weight), ctx.nbr), xv, gyr)
            if ctx.needs_input_grad[1]:
                wv = torch.zeros_like(weight).requires_grad_()
                (gw,) = torch.autograd.grad(_conv_fp32(round_tf32(x), wv, ctx.nbr), wv, gyr)
        return gx, gw, None


class _ConvBF16x2(torch.autograd.Function):
    """The same sum with split-bf16 operands in all three passes (see CONV_PRECISION)."""

    @staticmethod
    def forward(ctx, x, weight, nbr):
        ctx.save_for_backward(x, weight)
        ctx.nbr = nbr
        xh, xl = split_bf16(x)
        wh, wl = split_bf16(weight)
        if weight.shape[-2] <= 4:       # three-way weight split W = H + M + L; only xl*L is dropped
            w3, _ = split_bf16(weight - wh - wl)
            return _conv_fp32(xh + xl, wh + wl, nbr) + _conv_fp32(xh, w3, nbr)
        return _conv_fp32(xh + xl, wh, nbr) + _conv_fp32(xh, wl, nbr)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gh, gl = split_bf16(gy)
        xh, xl = split_bf16(x)
        wh, wl = split_bf16(weight)
        gx = gw = None
        with torch.enable_grad():
            if ctx.needs_input_grad[0]:
                xv = torch.zeros_like(x).requires_grad_()
                (g1,) = torch.autograd.grad(_conv_fp32(xv, wh, ctx.nbr), xv, gh + gl)
                xv = torch.zeros_like(x).requires_grad_()
                (g2,) = torch.autograd.grad(_conv_fp32(xv, wl, ctx.nbr), xv, gh)
                gx = g1 + g2
            if ctx.needs_input_grad[1]:
                wv = torch.zeros_like(weight).requires_grad_()
                if weight.shape[-2] <= 4:
                    (gw,) = torch.autograd.grad(_conv_fp32(xh + xl, wv, ctx.nbr), wv, gh + gl)
                else:
                    (g1,) = torch.autograd.grad(_conv_fp32(xh + xl, wv, ctx.nbr), wv, gh)
                    wv = torch.zeros_like(weight).requires_grad_()
                    (g2,) = torch.autograd.grad(_conv_fp32(xh, wv, ctx.nbr), wv, gl)
                    gw = g1 + g2
        return gx, gw, None


def conv(x: torch.Tensor, weight: torch.Tensor, nbr, bias: torch.Tensor | None = None) -> torch.Tensor:
    """MinkowskiConvolution forward: ``out[o] = bias + sum_k sum_{(i->o) in M_k} x[i] @ W[k]``.

    Call sites in the reference: ``modules/MinkowskiEngine/SENet.py:49-52,94-97``,
    ``resnet_block.py:48-54,95-107``, ``common.py:219-221``.  ``weight`` is ``[K^3, Cin, Cout]``
    (or ``[Cin, Cout]`` for the K=1, stride=1 ``use_mm`` case, where ``nbr`` is ignored)."""
    if weight.dim() != 2:
        nbr = _as_long(nbr)
    if CONV_PRECISION == "tf32" and x.dtype == torch.float32:
        out = _ConvTF32.apply(x, weight, nbr)
    elif CONV_PRECISION == "bf16x2" and x.dtype == torch.float32:
        out = _ConvBF16x2.apply(x, weight, nbr)
    else:
        out = _conv_fp32(x, weight, nbr)
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
