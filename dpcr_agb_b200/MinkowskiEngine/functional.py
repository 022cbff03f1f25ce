"""torch.autograd Functions over the C ABI -- one per MinkowskiEngine op on the MSENet hot path.

Every forward/backward here is a call into ``libb200sparse.so`` on the current CUDA stream; there is
no torch fallback.  Functions are declared fp32 (``custom_fwd(cast_inputs=float32)``) so a stray
half tensor under autocast never reaches the C ABI (SURVEY.md 8b "autocast interaction").
"""
from __future__ import annotations

import os

import torch
from torch.amp import custom_bwd, custom_fwd

from dpcr_agb_b200 import lib as L

# 0 = auto (tcgen05 where the shape qualifies, SIMT otherwise), 1 = force SIMT, 2 = force tcgen05
CONV_IMPL = int(os.environ.get("B2S_CONV_IMPL", "0"))

USE_PARITY_DGRAD = True   # tests flip this to compare the parity-plan dgrad with the dense transposed-table one

# bench.py sets this to a dict to collect ALGORITHMIC work per conv launch kind (pairs come from the neighbour
# table: one device reduction + host sync per launch, so only ever enabled in an untimed statistics pass)
WORK_STATS = None


def _account(kind, nbr, n_rows, c_in, c_out, k3, n_dev=None, kmap=None):
    """``kmap``: count the pairs through the map object (x-line maps never build ``nbr``)."""
    if WORK_STATS is None:
        return
    if n_dev is not None:
        n_rows = min(int(n_rows), int(n_dev.item()))
    if kmap is not None:
        pairs = kmap.num_pairs(n_rows)
    else:
        pairs = int((nbr[:, :n_rows] >= 0).sum().item()) if nbr is not None else int(n_rows)
    st = WORK_STATS.setdefault(kind, {"launches": 0, "pairs": 0, "flops": 0, "bytes": 0})
    st["launches"] += 1
    st["pairs"] += pairs
    st["flops"] += 2 * pairs * c_in * c_out
    # compulsory traffic: inputs + outputs + weights + int32 pair indices, each touched once
    st["bytes"] += 4 * (n_rows * (c_in + c_out)) + 4 * k3 * c_in * c_out + 8 * pairs


import contextlib  # noqa: E402

_NULL_CTX = contextlib.nullcontext()
_fwd = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


def round_tf32(x, n_dev=None):
    """TF32 round-to-nearest copy of a feature matrix (C ABI ``b2s_round_tf32``): the operand form the tensor-core
    kernels consume.  An operand used by several passes is rounded once and passed with ``prerounded=True``."""
    x = x.contiguous()
    y = torch.empty_like(x)
    if x.numel():
        L.call("b2s_round_tf32", x, x.shape[0], n_dev, x.numel() // x.shape[0], y)
    return y


# ---- TF32 twins -------------------------------------------------------------------------------------------------
# A producer whose result feeds a convolution can write the TF32-rounded operand itself (C ABI: the `*_tf32` output of
# b2s_maxpool_fwd / b2s_bn_apply / b2s_bn_bwd_apply / b2s_add_gelu_fwd).  The twin travels as a Python attribute of the
# plain tensor, stamped with that tensor's version counter; a convolution that finds a valid twin skips its own
# b2s_round_tf32 pass.  A tensor that IS the rounded result (no plain copy was written) is marked ``_b2s_is_tf32``.
# Nothing depends on the attribute surviving: without it the convolution rounds as before.
TWINS = True
FUSE_BIAS_GRAD = os.environ.get("B2S_FUSE_BIAS", "1") == "1"   # bn_bwd_apply also leaves the column sums of its gx
# A convolution's epilogue can accumulate the column sums / sums of squares of its output -- the statistics of the batch
# norm behind it (C ABI: col_stats of b2s_conv_gather_gemm / _lines_fwd).  OFF by default: measured on the B200 the
# transposing reduction in the (not overlapped) epilogue adds 0.27 ms to the convolutions of an MSENet14 step while the
# b2s_bn_stats passes it replaces cost 0.22 ms at 4 TB/s (profiles/r02_fusion_experiments.txt).
FUSE_BN_STATS = os.environ.get("B2S_FUSE_BN_STATS", "0") == "1"


def new_col_stats(n_out, c_out, device):
    """Buffer for the batch-norm statistics a convolution accumulates over its output (partial rows, see
    include/b200sparse.h ``col_stats``), or None when they are not wanted (no gradient mode: inference normalises with
    the running statistics)."""
    if not (FUSE_BN_STATS and torch.is_grad_enabled()) or n_out <= 0:
        return None
    return torch.empty(L.query("b2s_conv_col_stats_elems", n_out, c_out), dtype=torch.float32, device=device)


def attach_col_stats(y, stats):
    if stats is not None:
        y._b2s_colstats = (stats, y._version)
    return y


def col_stats_of(x, c):
    st = getattr(x, "_b2s_colstats", None)
    if st is not None and st[1] == x._version and st[0].numel() == L.query("b2s_conv_col_stats_elems", x.shape[0], c):
        return st[0]
    return None


def twins_on():
    return TWINS and CONV_IMPL != 1


def attach_twin(y, yr):
    y._b2s_tf32 = (yr, y._version)
    return y


def mark_rounded(y):
    y._b2s_is_tf32 = y._version
    return y


def rounded_operand(x, n_dev=None):
    """The TF32 operand form of ``x``: ``x`` itself if it was produced rounded, its twin if one is attached, else a
    fresh b2s_round_tf32 copy."""
    if getattr(x, "_b2s_is_tf32", None) == x._version:
        return x
    tw = getattr(x, "_b2s_tf32", None)
    if tw is not None and tw[1] == x._version and tw[0].shape == x.shape:
        return tw[0]
    return round_tf32(x, n_dev)


# ---- parameter gradients written in place ----------------------------------------------------------------------
# Inside ``with direct_param_grads():`` (dpcr_agb_b200.train: every parameter's ``.grad`` is a view of one flat,
# freshly zeroed gradient buffer) the backward kernels write weight / bias / batch-norm gradients straight into
# ``param.grad`` and return None to autograd, instead of returning a fresh tensor that AccumulateGrad then adds onto
# the zeros: one read-read-write pass over every parameter and ~50 launches per step less.
DIRECT_PARAM_GRADS = False


class direct_param_grads:
    def __enter__(self):
        global DIRECT_PARAM_GRADS
        self.old, DIRECT_PARAM_GRADS = DIRECT_PARAM_GRADS, True

    def __exit__(self, *exc):
        global DIRECT_PARAM_GRADS
        DIRECT_PARAM_GRADS = self.old


# ---- weight images built ahead -----------------------------------------------------------------------------------
# The tensor-core convolution kernels read the weights as a split, swizzled K-major image that every call used to build
# at the head of its workspace: 22 small launches per MSENet14 step on the critical path (0.2 ms).  Weights change once
# per optimiser step, so a trainer builds the images of all layers -- forward layout and dgrad layout -- at the start of
# the step on a side stream (``prepare_weight_images``); a convolution that finds a current image passes it as its
# workspace with w_layout bit 4 set and launches its main kernel only.  Without a current image nothing changes.
IMG_EPOCH = 0
_IMG_EVENTS = {}          # "fwd" / "bwd" -> event recorded behind the images of the current epoch


def conv_image_specs(model):
    """(kernel parameter, c_in, c_out, k3, dgrad w_layout) of every convolution of ``model`` whose shape the tensor-core
    kernels cover (the module class is looked up by name: this file must not import modules.py)."""
    specs = []
    for m in model.modules():
        if type(m).__name__ != "MinkowskiConvolution" or getattr(m, "is_transpose", False):
            continue
        c_in, c_out, k3 = m.in_channels, m.out_channels, m.kernel_volume
        if L.query("b2s_conv_weight_image_bytes", c_in, c_out, k3) <= 0 or \
                L.query("b2s_conv_weight_image_bytes", c_out, c_in, k3) <= 0:
            continue
        symmetric = k3 > 1 and all(s == 1 for s in m.stride) and all(k % 2 == 1 for k in m.kernel_size)
        specs.append((m.kernel, c_in, c_out, k3, 3 if symmetric else 1))
    return specs


def prepare_weight_images(specs, stream):
    """Build the forward and dgrad weight images of ``specs`` on ``stream`` (forked from the current stream here)."""
    global IMG_EPOCH
    IMG_EPOCH += 1
    cur = torch.cuda.current_stream()
    stream.wait_stream(cur)
    with torch.cuda.stream(stream):
        for phase in ("fwd", "bwd"):
            for kernel, c_in, c_out, k3, dl in specs:
                layout, ci, co = (0, c_in, c_out) if phase == "fwd" else (dl, c_out, c_in)
                imgs = kernel.__dict__.setdefault("_b2s_img", {})
                ent = imgs.get(layout)
                if ent is None:
                    nbytes = L.query("b2s_conv_weight_image_bytes", ci, co, k3)
                    ent = imgs[layout] = [torch.empty(nbytes, dtype=torch.uint8, device=kernel.device), -1]
                L.call("b2s_conv_weight_image", kernel, ci, co, k3, layout, ent[0], ent[0].numel())
                ent[1] = IMG_EPOCH
            ev = torch.cuda.Event()
            ev.record(stream)
            _IMG_EVENTS[phase] = ev


def weight_image(kernel, layout):
    """The current prebuilt image of ``kernel`` in ``layout`` (the caller's stream is made to wait for it), or None."""
    ent = getattr(kernel, "_b2s_img", {}).get(layout)
    if ent is None or ent[1] != IMG_EPOCH:
        return None
    ev = _IMG_EVENTS.get("fwd" if layout == 0 else "bwd")
    if ev is not None:
        torch.cuda.current_stream().wait_event(ev)
    return ent[0]


# ---- weight gradients beside the dgrad chain ---------------------------------------------------------------------
# Inside ``with side_wgrad(stream):`` (and direct_param_grads) the weight-gradient kernels are launched on ``stream``:
# they are leaves of the backward pass -- nothing but the optimiser waits for them -- while the chain
# bn_bwd -> dgrad -> bn_bwd -> ... is the critical path, and the two use different resources (wgrad: shared-memory
# bandwidth, batch norm: HBM).  Every launch waits for the main stream's position (its operands are ready), the context
# exit makes the main stream wait for the side stream.
WGRAD_STREAM = None


class side_wgrad:
    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        global WGRAD_STREAM
        self.old, WGRAD_STREAM = WGRAD_STREAM, self.stream

    def __exit__(self, *exc):
        global WGRAD_STREAM
        WGRAD_STREAM = self.old
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def _direct(p):
    return (DIRECT_PARAM_GRADS and p is not None and p.grad is not None and p.grad.is_contiguous()
            and p.grad.dtype == torch.float32 and p.grad.shape == p.shape)


def _claim(p):
    """A backward kernel is about to OVERWRITE ``p.grad`` in place (direct mode).  A second write before the next
    ``zero_grad`` -- a second ``backward()`` for gradient accumulation, or a parameter shared by two ops -- would
    silently discard the first gradient, so it raises instead (``FlatAdaBelief.zero_grad`` clears the marks)."""
    if getattr(p, "_b2s_grad_written", False):
        raise RuntimeError("direct_param_grads(): this parameter's gradient was already written in place since the "
                           "last zero_grad(); accumulate outside direct_param_grads() instead")
    p._b2s_grad_written = True


def _ws(n_in, n_out, c_in, c_out, k3, device, prerounded=False):
    nbytes = L.query("b2s_conv_workspace_bytes", n_in, n_out, c_in, c_out, k3, 1 if prerounded else 0)
    if nbytes < 0:
        raise L.B2SError("b2s_conv_workspace_bytes rejected the shape")
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device), nbytes


def dgrad_strided(gy, w, kmap, c_gy, c_x, wimg=None):
    """gx[i] = sum_k gy[inv[k,i]] @ W[k]^T for a stride-2 map through the parity plan (C ABI
    ``b2s_conv_dgrad_strided``); ``gy`` must be TF32-rounded.  ``wimg``: prebuilt weight image (layout 1)."""
    perm, bounds = kmap.parity_plan
    gx = torch.empty((kmap.n_in, c_x), dtype=torch.float32, device=gy.device)
    _account("dgrad", kmap.inv, kmap.n_in, c_gy, c_x, kmap.k3, kmap.n_in_dev)
    nbytes = L.query("b2s_conv_dgrad_strided_workspace_bytes", c_gy, c_x, kmap.k3)
    if wimg is not None and wimg.numel() >= nbytes:
        ws, flags = wimg, 1
    else:
        ws, flags = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=gy.device), 0
    L.call("b2s_conv_dgrad_strided", gy, w, kmap.inv, perm, bounds, kmap.n_in, kmap.n_in_dev, c_gy, c_x,
           L.host_i32(*kmap.kernel_size), gx, ws, ws.numel() if flags else nbytes, flags)
    return gx


def _tc(impl):
    return (CONV_IMPL if impl is None else impl) != 1


# Shapes the tensor-core kernels cover -- mirrors b2s_conv_tc_supported / b2s_wgrad_tc_supported (csrc/conv_tc.cu,
# wgrad_tc.cu).  Only these may be handed operands in operand form (``prerounded``): in the default split-bf16 mode
# that form is a packed bit pattern, not a rounded float, and the library rejects it on the SIMT path.
def _fwd_tc_ok(c_in, c_out):
    return c_out % 64 == 0 and (c_in <= 4 or c_in % 32 == 0)


def _wg_tc_ok(c_in, c_out, has_map):
    if c_in <= 4:
        return has_map and c_out % 32 == 0
    blocks = c_in // 32
    return c_in % 32 == 0 and c_out % 64 == 0 and (blocks >= 8 or (blocks & (blocks - 1)) == 0)


def gather_gemm(x, w, bias, nbr, n_in, n_out, c_in, c_out, k3, w_layout, impl=None, n_out_dev=None,
                prerounded=False, col_stats=None, wimg=None):
    """y[o] = bias + sum_k x[nbr[k,o]] @ B_k  (C ABI ``b2s_conv_gather_gemm``).  ``n_out_dev``: device row count
    (then ``n_out`` is the capacity / pitch of ``nbr``).  ``prerounded``: x is already TF32-representable.
    ``col_stats``: float64 [2 c_out + 1] that receives the column sums / sums of squares of y."""
    y = torch.empty((n_out, c_out), dtype=torch.float32, device=x.device)
    _account("dgrad" if (w_layout & 1) else "fwd", nbr, n_out, c_in, c_out, k3, n_out_dev)
    if wimg is not None and prerounded and c_in > 4 and (CONV_IMPL if impl is None else impl) != 1:
        # ``wimg``: the prebuilt weight image of (w, w_layout) -- it IS the workspace of a pre-rounded call
        L.call("b2s_conv_gather_gemm", x, w, bias, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3, w_layout | 4 | 16, y,
               wimg, wimg.numel(), CONV_IMPL if impl is None else impl, col_stats)
        return y
    ws, nbytes = _ws(n_in, n_out, c_in, c_out, k3, x.device, prerounded)
    L.call("b2s_conv_gather_gemm", x, w, bias, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3,
           w_layout | (4 if prerounded else 0), y, ws, nbytes, CONV_IMPL if impl is None else impl, col_stats)
    return y


def wgrad(x, gy, nbr, n_in, n_out, c_in, c_out, k3, impl=None, n_out_dev=None, prerounded=False, out=None):
    """``out``: contiguous fp32 storage of k3*c_in*c_out elements to write into (every element is written)."""
    gw = out.view(k3, c_in, c_out) if out is not None else torch.empty((k3, c_in, c_out), dtype=torch.float32,
                                                                      device=x.device)
    _account("wgrad", nbr, n_out, c_in, c_out, k3, n_out_dev)
    ws, nbytes = _ws(n_in, n_out, c_in, c_out, k3, x.device, prerounded)
    L.call("b2s_conv_wgrad", x, gy, nbr, n_in, n_out, n_out_dev, c_in, c_out, k3, gw, ws, nbytes,
           CONV_IMPL if impl is None else impl, 1 if prerounded else 0)
    return gw


def lines_path(kmap, c_in, c_out):
    """True when the convolution runs through the x-line form of its map (C ABI ``b2s_conv_lines_*``: few input
    channels on the quantiser's rows, split-bf16 mode -- the k7 stem)."""
    return (kmap is not None and _tc(None) and c_in <= 4 and kmap.lines_ok
            and L.query("b2s_conv_lines_supported", c_in, c_out, L.host_i32(*kmap.kernel_size)) == 1)


def lines_fwd(x, w, bias, kmap, c_in, c_out, col_stats=None):
    y = torch.empty((kmap.n_out, c_out), dtype=torch.float32, device=x.device)
    _account("fwd", None, kmap.n_out, c_in, c_out, kmap.k3, kmap.n_out_dev, kmap=kmap)
    ks = L.host_i32(*kmap.kernel_size)
    nbytes = L.query("b2s_conv_lines_workspace_bytes", kmap.n_in, c_in, c_out, ks)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
    L.call("b2s_conv_lines_fwd", x, w, bias, kmap.lines, kmap.n_in, kmap.n_out, kmap.n_out_dev, c_in, c_out, ks, y, ws,
           nbytes, col_stats)
    return y


def lines_wgrad(x, gy_operand, kmap, c_in, c_out, out=None):
    gw = out.view(kmap.k3, c_in, c_out) if out is not None else torch.empty((kmap.k3, c_in, c_out), dtype=torch.float32,
                                                                           device=x.device)
    _account("wgrad", None, kmap.n_out, c_in, c_out, kmap.k3, kmap.n_out_dev, kmap=kmap)
    ks = L.host_i32(*kmap.kernel_size)
    nbytes = L.query("b2s_conv_lines_workspace_bytes", kmap.n_in, c_in, c_out, ks)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=x.device)
    L.call("b2s_conv_lines_wgrad", x, gy_operand, kmap.lines, kmap.n_in, kmap.n_out, kmap.n_out_dev, c_in, c_out, ks,
           gw, ws, nbytes)
    return gw


class ConvolutionFunction(torch.autograd.Function):
    """MinkowskiConvolution fwd / dgrad / wgrad (reference call sites: SENet.py:49-52,94-97;
    resnet_block.py:48-54,95-107).  ``kmap`` is None for the K=1, stride=1 ``use_mm`` case."""

    @staticmethod
    @_fwd
    def forward(ctx, feats, kernel, bias, kmap, n_dev=None, col_stats=None):
        """``n_dev``: device row count of the (identity-map) input when ``kmap`` is None; a KernelMap carries its
        own device counts.  ``col_stats``: see :func:`gather_gemm` (filled as a side effect, not differentiable)."""
        feats = feats.contiguous()
        kernel = kernel.contiguous()
        c_in, c_out = kernel.shape[-2], kernel.shape[-1]
        use_lines = lines_path(kmap, c_in, c_out)
        if kmap is None:
            n_in = n_out = feats.shape[0]
            nbr, k3 = None, 1
            nd_in = nd_out = n_dev
        else:
            n_in, n_out, k3 = kmap.n_in, kmap.n_out, kmap.k3
            nbr = None if use_lines else kmap.nbr
            nd_in, nd_out = kmap.n_in_dev, kmap.n_out_dev
        assert feats.shape == (n_in, c_in), f"feature shape {tuple(feats.shape)} does not match the map ({n_in},{c_in})"
        b = bias.contiguous().view(-1) if bias is not None else None
        ctx.lines = use_lines
        if use_lines:
            out = lines_fwd(feats, kernel, b, kmap, c_in, c_out, col_stats)
            ctx.pre, ctx.kmap, ctx.nd, ctx.dims = False, kmap, (nd_in, nd_out), (n_in, n_out, c_in, c_out, k3)
            ctx.has_bias, ctx.params = bias is not None, (kernel, bias)
            ctx.save_for_backward(feats, kernel)
            return out
        # the tensor-core kernels consume TF32 operands: x is rounded once here and the rounded copy is what is
        # saved for wgrad (c_in <= 4: the stem pads + rounds inside the library)
        pre = _tc(None) and c_in > 4 and _fwd_tc_ok(c_in, c_out)
        raw = feats
        if pre:
            feats = rounded_operand(feats, nd_in)
        out = gather_gemm(feats, kernel, b, nbr, n_in, n_out, c_in, c_out, k3, 0, n_out_dev=nd_out, prerounded=pre,
                          col_stats=col_stats, wimg=weight_image(kernel, 0) if pre else None)
        if pre and not _wg_tc_ok(c_in, c_out, kmap is not None):   # wgrad will run on the SIMT kernel: keep plain x
            feats, pre = raw, False
        ctx.pre = pre
        ctx.kmap = kmap
        ctx.nd = (nd_in, nd_out)
        ctx.dims = (n_in, n_out, c_in, c_out, k3)
        ctx.has_bias = bias is not None
        ctx.params = (kernel, bias)
        ctx.save_for_backward(feats, kernel)
        return out

    @staticmethod
    @_bwd
    def backward(ctx, gy):
        feats, kernel = ctx.saved_tensors
        kmap = ctx.kmap
        n_in, n_out, c_in, c_out, k3 = ctx.dims
        gy = gy.contiguous()
        nd_in, nd_out = ctx.nd
        gx = gw = gb = None
        # grad_out feeds dgrad and wgrad: brought into operand form once, for whichever of the two runs on tensor cores
        dg_ok = _tc(None) and c_out > 4 and _fwd_tc_ok(c_out, c_in) and ctx.needs_input_grad[0]
        wg_ok = (_tc(None) and c_out % 32 == 0 and _wg_tc_ok(c_in, c_out, kmap is not None)
                 and (ctx.pre or c_in <= 4) and ctx.needs_input_grad[1])
        gyr = rounded_operand(gy, nd_out) if (dg_ok or wg_ok) else gy
        if ctx.needs_input_grad[0]:
            gd, pre_gy = (gyr, True) if dg_ok else (gy, False)
            dl = 3 if (kmap is not None and kmap.symmetric) else 1
            img = weight_image(ctx.params[0], dl) if pre_gy else None
            if kmap is None:
                gx = gather_gemm(gd, kernel, None, None, n_out, n_in, c_out, c_in, 1, 1, n_out_dev=nd_in,
                                 prerounded=pre_gy, wimg=img)
            elif kmap.symmetric:      # transposed map == same table with the kernel index reversed
                gx = gather_gemm(gd, kernel, None, kmap.nbr, n_out, n_in, c_out, c_in, k3, 1 | 2, n_out_dev=nd_in,
                                 prerounded=pre_gy, wimg=img)
            elif (USE_PARITY_DGRAD and pre_gy and CONV_IMPL == 0 and c_out % 32 == 0 and c_in % 64 == 0
                  and kmap.parity_plan is not None):
                gx = dgrad_strided(gd, kernel, kmap, c_out, c_in, wimg=img)
            else:
                gx = gather_gemm(gd, kernel, None, kmap.inv, n_out, n_in, c_out, c_in, k3, 1, n_out_dev=nd_in,
                                 prerounded=pre_gy, wimg=img)
        if ctx.needs_input_grad[1]:
            both = wg_ok                                  # feats is the operand-form copy saved by forward
            kp = ctx.params[0]
            g_op = gyr if both else gy
            side = WGRAD_STREAM if (_direct(kp) and WGRAD_STREAM is not None and WORK_STATS is None) else None
            if side is not None:                          # in-place gradient on the side stream (see side_wgrad)
                side.wait_stream(torch.cuda.current_stream())
                feats.record_stream(side)
                g_op.record_stream(side)
            with (torch.cuda.stream(side) if side is not None else _NULL_CTX):
                if ctx.lines and wg_ok:
                    gw = lines_wgrad(feats, gyr, kmap, c_in, c_out,
                                     out=kp.grad if _direct(kp) else None).view(kernel.shape)
                else:
                    gw = wgrad(feats, g_op, None if kmap is None else kmap.nbr, n_in, n_out, c_in, c_out,
                               k3, n_out_dev=nd_out, prerounded=both,
                               out=kp.grad if _direct(kp) else None).view(kernel.shape)
            if _direct(kp):
                _claim(kp)
                gw = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            bp = ctx.params[1]
            pre = getattr(gy, "_b2s_colsum", None)      # column sums already accumulated by the producer of gy
            if pre is not None and pre[1] == gy._version and pre[0].dim() == 2 and pre[0].shape[1] == c_out:
                gb = bp.grad.view(1, c_out) if _direct(bp) else torch.empty((1, c_out), dtype=torch.float32,
                                                                             device=gy.device)
                L.call("b2s_sum_rows", pre[0], pre[0].shape[0], c_out, gb)
                if _direct(bp):
                    _claim(bp)
                    gb = None
            else:
                gb = bp.grad.view(1, c_out) if _direct(bp) else torch.empty((1, c_out), dtype=torch.float32,
                                                                             device=gy.device)
                L.call("b2s_colsum", gy, n_out, nd_out, c_out, gb)
                if _direct(bp):
                    _claim(bp)
                    gb = None
        return gx, gw, gb, None, None, None


class MaxPoolFunction(torch.autograd.Function):
    """MinkowskiMaxPooling (SENet.py:53)."""

    @staticmethod
    @_fwd
    def forward(ctx, feats, kmap, twin=False):
        feats = feats.contiguous()
        c = feats.shape[1]
        y = torch.empty((kmap.n_out, c), dtype=torch.float32, device=feats.device)
        arg = torch.empty((kmap.n_out, c), dtype=torch.int32, device=feats.device)
        yr = torch.empty_like(y) if twin else None
        L.call("b2s_maxpool_fwd", feats, kmap.nbr, kmap.n_out, kmap.n_out_dev, c, kmap.k3, y, arg, yr)
        ctx.save_for_backward(arg)
        ctx.nd_out = kmap.n_out_dev
        ctx.dims = (kmap.n_in, kmap.n_out, c)
        if twin:
            ctx.mark_non_differentiable(yr)
            return y, yr
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy, *unused):
        (arg,) = ctx.saved_tensors
        n_in, n_out, c = ctx.dims
        gx = torch.empty((n_in, c), dtype=torch.float32, device=gy.device)
        L.call("b2s_maxpool_bwd", gy.contiguous(), arg, n_in, n_out, ctx.nd_out, c, gx)
        return gx, None, None


class LocalPoolFunction(torch.autograd.Function):
    """MinkowskiSumPooling / MinkowskiAvgPooling (networks.py:29): ``y[o] = scale[o] * sum_k x[nbr[k, o]]`` with
    ``scale = 1 / #inputs under the kernel`` for the average (C ABI ``b2s_sumpool`` / ``b2s_nbr_inv_counts``);
    backward = the same kernel on the transposed table with the scale applied per gathered row."""

    @staticmethod
    @_fwd
    def forward(ctx, feats, kmap, average):
        feats = feats.contiguous()
        c = feats.shape[1]
        y = torch.empty((kmap.n_out, c), dtype=torch.float32, device=feats.device)
        inv = None
        if average:
            inv = torch.empty(kmap.n_out, dtype=torch.float32, device=feats.device)
            L.call("b2s_nbr_inv_counts", kmap.nbr, kmap.k3, kmap.n_out, kmap.n_out_dev, inv)
        L.call("b2s_sumpool", feats, kmap.nbr, None, inv, kmap.n_out, kmap.n_out_dev, c, kmap.k3, y)
        ctx.kmap, ctx.inv, ctx.c = kmap, inv, c
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy):
        kmap = ctx.kmap
        gx = torch.empty((kmap.n_in, ctx.c), dtype=torch.float32, device=gy.device)
        L.call("b2s_sumpool", gy.contiguous(), kmap.inv, ctx.inv, None, kmap.n_in, kmap.n_in_dev, ctx.c, kmap.k3, gx)
        return gx, None, None


class GlobalPoolFunction(torch.autograd.Function):
    """MinkowskiGlobal{Sum,Avg}Pooling / MinkowskiGlobalPooling (senet_block.py:43; common.py:44-48)."""

    @staticmethod
    @_fwd
    def forward(ctx, feats, coords, num_batches, scale, n_dev=None):
        feats = feats.contiguous()
        n, c = feats.shape
        y = torch.empty((num_batches, c), dtype=torch.float32, device=feats.device)
        L.call("b2s_segment_sum", feats, coords, 4, n, n_dev, c, num_batches, scale, y)
        ctx.coords, ctx.scale, ctx.dims, ctx.nd = coords, scale, (n, c), n_dev
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy):
        n, c = ctx.dims
        gx = torch.empty((n, c), dtype=torch.float32, device=gy.device)
        L.call("b2s_segment_bcast", gy.contiguous(), ctx.coords, 4, n, ctx.nd, c, ctx.scale, gx)
        return gx, None, None, None, None


class GlobalMaxPoolFunction(torch.autograd.Function):
    """MinkowskiGlobalMaxPooling (networks.py:39; PointNet.py:28 via GLOBAL_POOL["max"])."""

    @staticmethod
    @_fwd
    def forward(ctx, feats, coords, num_batches, n_dev=None):
        feats = feats.contiguous()
        n, c = feats.shape
        y = torch.empty((num_batches, c), dtype=torch.float32, device=feats.device)
        arg = torch.empty((num_batches, c), dtype=torch.int32, device=feats.device)
        L.call("b2s_segment_max", feats, coords, 4, n, n_dev, c, num_batches, y, arg)
        ctx.save_for_backward(arg)
        ctx.dims = (n, c, num_batches)
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy):
        (arg,) = ctx.saved_tensors
        n, c, nb = ctx.dims
        gx = torch.empty((n, c), dtype=torch.float32, device=gy.device)
        L.call("b2s_segment_max_bwd", gy.contiguous(), arg, n, c, nb, gx)
        return gx, None, None, None


class BroadcastFunction(torch.autograd.Function):
    """MinkowskiBroadcast / MinkowskiBroadcastAddition: ``out[i] = (x[i] +) y[batch(i)]``; backward of the broadcast
    operand = per-plot sum."""

    @staticmethod
    @_fwd
    def forward(ctx, y, coords, n, n_dev=None):
        y = y.contiguous()
        out = torch.empty((n, y.shape[1]), dtype=torch.float32, device=y.device)
        L.call("b2s_segment_bcast", y, coords, 4, n, n_dev, y.shape[1], None, out)
        ctx.coords, ctx.nd, ctx.nb = coords, n_dev, y.shape[0]
        return out

    @staticmethod
    @_bwd
    def backward(ctx, g):
        g = g.contiguous()
        gy = torch.empty((ctx.nb, g.shape[1]), dtype=torch.float32, device=g.device)
        L.call("b2s_segment_sum", g, ctx.coords, 4, g.shape[0], ctx.nd, g.shape[1], ctx.nb, None, gy)
        return gy, None, None, None


class BroadcastMulFunction(torch.autograd.Function):
    """MinkowskiBroadcastMultiplication (senet_block.py:44,50): out[i] = x[i] * y[batch(i)]."""

    @staticmethod
    @_fwd
    def forward(ctx, x, y, coords, num_batches, n_dev=None):
        x, y = x.contiguous(), y.contiguous()
        n, c = x.shape
        assert y.shape[0] == num_batches and y.shape[1] in (1, c)
        out = torch.empty_like(x)
        L.call("b2s_bcast_mul_fwd", x, y, coords, 4, n, n_dev, c, y.shape[1], out)
        ctx.save_for_backward(x, y)
        ctx.coords, ctx.nb, ctx.nd = coords, num_batches, n_dev
        return out

    @staticmethod
    @_bwd
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        n, c = x.shape
        g = g.contiguous()
        need_x, need_y = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if y.shape[1] != c:                      # per-plot scalar (drop-path mask): no gradient to the mask
            gx = torch.empty_like(x)
            L.call("b2s_bcast_mul_fwd", g, y, ctx.coords, 4, n, ctx.nd, c, 1, gx)
            return gx, None, None, None, None
        gx = torch.empty_like(x) if need_x else None
        gy = torch.empty_like(y) if need_y else None
        L.call("b2s_bcast_mul_bwd", g, x, y, ctx.coords, 4, n, ctx.nd, c, ctx.nb, gx, gy)
        return gx, gy, None, None, None


class BatchNormFunction(torch.autograd.Function):
    """MinkowskiBatchNorm == nn.BatchNorm1d over all rows (SENet.py:35,51,98); act=1 fuses exact GELU."""

    @staticmethod
    @_fwd
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, act, n_dev=None,
                tf32_only=False, col_stats=None):
        """``tf32_only``: the result is written TF32-rounded and nothing else (its only consumer is a convolution).
        ``col_stats``: column sums / sums of squares of x already accumulated by the convolution that produced it."""
        x = x.contiguous()
        n, c = x.shape
        dev = x.device
        if training:
            mean = torch.empty(c, dtype=torch.float32, device=dev)
            invstd = torch.empty(c, dtype=torch.float32, device=dev)
            if col_stats is not None and n > 0:
                L.call("b2s_bn_finalize", col_stats, n, n_dev, c, float(eps), float(momentum), running_mean,
                       running_var, mean, invstd)
            else:
                ws = torch.empty(2 * c + 1, dtype=torch.float64, device=dev)
                L.call("b2s_bn_stats", x, n, n_dev, c, float(eps), float(momentum), running_mean, running_var, ws,
                       mean, invstd)
        else:
            mean = running_mean
            invstd = torch.rsqrt(running_var + eps)
        y = torch.empty_like(x)
        if tf32_only:
            L.call("b2s_bn_apply", x, mean, invstd, weight, bias, n, n_dev, c, act, None, y)
        else:
            L.call("b2s_bn_apply", x, mean, invstd, weight, bias, n, n_dev, c, act, y, None)
        ctx.save_for_backward(x, mean, invstd, weight, bias)
        ctx.training, ctx.act, ctx.nd = training, act, n_dev
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy):
        x, mean, invstd, weight, bias = ctx.saved_tensors
        n, c = x.shape
        gy = gy.contiguous()
        dev = x.device
        sums = torch.empty(2 * c, dtype=torch.float32, device=dev)
        ws = torch.empty(2 * c + 1, dtype=torch.float64, device=dev)
        L.call("b2s_bn_bwd_reduce", gy, x, mean, invstd, weight, bias, n, ctx.nd, c, ctx.act, ws, sums)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            # the input gradient of a batch norm is the output gradient of the convolution in front of it: write
            # the TF32 operand for its dgrad / wgrad alongside
            gxr = torch.empty_like(x) if (twins_on() and c > 4) else None
            # ... and its column sums are that convolution's bias gradient: accumulated here while gx is written
            cs = None
            if FUSE_BIAS_GRAD:          # one partial row per block of the launch; the consumer adds them
                cs = torch.empty((L.query("b2s_bn_bwd_colsum_rows", n, c), c), dtype=torch.float32, device=dev)
            L.call("b2s_bn_bwd_apply", gy, x, mean, invstd, weight, bias, sums, n, ctx.nd, c, ctx.act,
                   1 if ctx.training else 0, gx, gxr, cs)
            if gxr is not None:
                attach_twin(gx, gxr)
            if cs is not None:
                gx._b2s_colsum = (cs, gx._version)
        gw = gb = None
        if weight is not None and ctx.needs_input_grad[1]:
            if _direct(weight):
                _claim(weight)
                weight.grad.copy_(sums[c:])
            else:
                gw = sums[c:].clone()
        if bias is not None and ctx.needs_input_grad[2]:
            if _direct(bias):
                _claim(bias)
                bias.grad.copy_(sums[:c])
            else:
                gb = sums[:c].clone()
        return gx, gw, gb, None, None, None, None, None, None, None, None, None


class GELUFunction(torch.autograd.Function):
    """NL.MinkowskiGELU: exact-erf GELU (common.py:41)."""

    @staticmethod
    @_fwd
    def forward(ctx, x, n_dev=None):
        x = x.contiguous()
        y = torch.empty_like(x)
        n = x.shape[0] if x.dim() > 1 else 1
        L.call("b2s_gelu_fwd", x, n, n_dev if x.dim() > 1 else None, x.numel() // max(n, 1), y)
        ctx.save_for_backward(x)
        ctx.nd = n_dev if x.dim() > 1 else None
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        n = x.shape[0] if x.dim() > 1 else 1
        L.call("b2s_gelu_bwd", gy.contiguous(), x, n, ctx.nd, x.numel() // max(n, 1), gx)
        return gx, None


class AddGELUFunction(torch.autograd.Function):
    """Residual join ``act(a + b)`` of every block (senet_block.py:93-94) as one kernel; the sum is kept for the
    backward, which is a single GELU-gradient kernel shared by both addends."""

    @staticmethod
    @_fwd
    def forward(ctx, a, b, n_dev=None, tf32_out=0):
        """``tf32_out``: 0 plain result; 1 plain result + TF32 twin (returned second); 2 TF32-rounded result only."""
        a, b = a.contiguous(), b.contiguous()
        s, y = torch.empty_like(a), torch.empty_like(a)
        n = a.shape[0]
        c = a.numel() // max(n, 1)
        ctx.save_for_backward(s)
        ctx.nd = n_dev
        if tf32_out == 1:
            yr = torch.empty_like(a)
            L.call("b2s_add_gelu_fwd", a, b, n, n_dev, c, s, y, yr)
            ctx.mark_non_differentiable(yr)
            return y, yr
        if tf32_out == 2:
            L.call("b2s_add_gelu_fwd", a, b, n, n_dev, c, s, None, y)
        else:
            L.call("b2s_add_gelu_fwd", a, b, n, n_dev, c, s, y, None)
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy, *unused):
        (s,) = ctx.saved_tensors
        g = torch.empty_like(s)
        n = s.shape[0]
        L.call("b2s_gelu_bwd", gy.contiguous(), s, n, ctx.nd, s.numel() // max(n, 1), g)
        return g, g, None, None


class SETailFunction(torch.autograd.Function):
    """The tail of an SE residual block as one op (senet_block.py:33-50 SELayer, :83-94 drop path + residual + act):
    ``y = gelu(u * (sigmoid(fc2(gelu(fc1(mean_plot(u))))) * keep[plot]) + res)``.  Forward = per-plot mean, one MLP
    launch for all plots, one gated add+GELU pass over the rows; backward = one pass over the rows (residual
    gradient, gated gradient, per-plot product sums), one MLP launch pair, one broadcast add of the pooled branch."""

    @staticmethod
    @_fwd
    def forward(ctx, u, res, w1, b1, w2, b2, keep, coords, num_batches, inv_counts, n_dev=None, tf32_out=0):
        u, res = u.contiguous(), res.contiguous()
        n, c = u.shape
        h = w1.shape[0]
        dev = u.device
        w1c, w2c = w1.contiguous(), w2.contiguous()
        keep = keep.contiguous().view(-1) if keep is not None else None
        pooled = torch.empty((num_batches, c), dtype=torch.float32, device=dev)
        L.call("b2s_segment_sum", u, coords, 4, n, n_dev, c, num_batches, inv_counts, pooled)
        h_pre = torch.empty((num_batches, h), dtype=torch.float32, device=dev)
        gate, gate_eff = torch.empty_like(pooled), torch.empty_like(pooled)
        L.call("b2s_se_gate_fwd", pooled, w1c, b1, w2c, b2, keep, num_batches, c, h, h_pre, gate, gate_eff)
        s, y = torch.empty_like(u), torch.empty_like(u)
        yr = torch.empty_like(u) if tf32_out == 1 else None
        if tf32_out == 2:
            L.call("b2s_gated_add_gelu_fwd", u, gate_eff, res, coords, 4, n, n_dev, c, s, None, y)
        else:
            L.call("b2s_gated_add_gelu_fwd", u, gate_eff, res, coords, 4, n, n_dev, c, s, y, yr)
        ctx.save_for_backward(u, s, pooled, h_pre, gate, gate_eff, w1c, w2c, keep, inv_counts)
        ctx.coords, ctx.nb, ctx.nd = coords, num_batches, n_dev
        ctx.params = (w1, b1, w2, b2)
        if yr is not None:
            ctx.mark_non_differentiable(yr)
            return y, yr
        return y

    @staticmethod
    @_bwd
    def backward(ctx, gy, *unused):
        u, s, pooled, h_pre, gate, gate_eff, w1c, w2c, keep, inv_counts = ctx.saved_tensors
        n, c = u.shape
        h, nb, dev = w1c.shape[0], ctx.nb, u.device
        g_res, g_u = torch.empty_like(u), torch.empty_like(u)
        g_ge = torch.empty((nb, c), dtype=torch.float32, device=dev)
        L.call("b2s_gated_add_gelu_bwd", gy.contiguous(), s, u, gate_eff, ctx.coords, 4, n, ctx.nd, c, nb, g_res, g_u,
               g_ge)
        w1, b1, w2, b2 = ctx.params
        outs = []
        for prm in (w1, b1, w2, b2):
            if prm is None:
                outs.append(None)
            elif _direct(prm):
                _claim(prm)
                outs.append(prm.grad)
            else:
                outs.append(torch.empty_like(prm, memory_format=torch.contiguous_format))
        gz2 = torch.empty((nb, c), dtype=torch.float32, device=dev)
        ghp = torch.empty((2, nb, h), dtype=torch.float32, device=dev)
        g_pooled = torch.empty((nb, c), dtype=torch.float32, device=dev)
        L.call("b2s_se_gate_bwd", g_ge, keep, gate, h_pre, pooled, w1c, w2c, inv_counts, nb, c, h, gz2, ghp, g_pooled,
               outs[0], outs[1], outs[2], outs[3])
        L.call("b2s_bcast_add_", g_u, g_pooled, ctx.coords, 4, n, ctx.nd, c)
        grads = [None if (prm is None or _direct(prm)) else o for prm, o in zip((w1, b1, w2, b2), outs)]
        return (g_u, g_res, grads[0], grads[1], grads[2], grads[3], None, None, None, None, None, None)
