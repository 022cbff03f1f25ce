"""GPU parity of the convolution kernels (a6-a8) against the oracle, SIMT and tcgen05 paths, through
the C ABI.  The tcgen05 kernels run in both operand modes (``b2s_set_tuning("precise", ...)``): the default
split-bf16 mode ("bf16x2": 16-17 significant bits per operand) and the TF32 mode.  Two bars each: against the
fp32/fp64 oracle -- 1e-3 relative (max|a-b| / max|b|), the fp32/TF32 tolerance of BASELINE.json, for TF32 and
5e-5 for split-bf16 -- and MODEL_TOL against the oracle's precision model of the same mode (oracle/ops.py
CONV_PRECISION), which restates exactly what the kernels compute and leaves only the summation order free."""
import numpy as np
import pytest
import torch

from dpcr_agb_b200 import lib as L
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from oracle import coords as oc
from oracle import ops as oo
import b2s_testutil as util

pytestmark = pytest.mark.gpu

SIMT, TC = 1, 2
TF32_MODEL_TOL = 2e-5
FP32_TOL = {"tf32": 1e-3, "bf16x2": 5e-5}     # tcgen05 kernels against the fp32 / fp64 oracle, per operand mode
_MODEL = ["bf16x2"]


@pytest.fixture(params=["bf16x2", "tf32"], autouse=True)
def opmode(request):
    """Runs every test of this file in both operand modes of the tensor-core kernels."""
    _MODEL[0] = request.param
    L.set_tuning("precise", 1 if request.param == "bf16x2" else 0)
    yield request.param
    L.set_tuning("precise", -1)


class tf32_model:
    """Context manager: evaluate the oracle convolution under the precision model of the current operand mode."""

    def __enter__(self):
        self.old, oo.CONV_PRECISION = oo.CONV_PRECISION, _MODEL[0]

    def __exit__(self, *a):
        oo.CONV_PRECISION = self.old


def _tol(impl):
    return FP32_TOL[_MODEL[0]] if impl != SIMT else 2e-5


def _maps(n, nb=2, extent=9, seed=0, K=3, strided=False):
    rng = np.random.default_rng(seed)
    c = util.random_coords(rng, n, nb=nb, extent=extent)
    if strided:
        out, _ = oc.stride_map(c, (2, 2, 2))
    else:
        out = c
    nbr = oc.kernel_map_table(c, out, K, (1, 1, 1))
    return c, out, nbr


def _run_fwd(x, w, b, nbr, n_in, n_out, impl, dev):
    xg, wg = torch.from_numpy(x).to(dev), torch.from_numpy(w).to(dev)
    bg = torch.from_numpy(b).to(dev) if b is not None else None
    ng = torch.from_numpy(nbr).to(dev) if nbr is not None else None
    k3 = 1 if nbr is None else nbr.shape[0]
    return Fn.gather_gemm(xg, wg, bg, ng, n_in, n_out, w.shape[-2], w.shape[-1], k3, 0, impl=impl)


@pytest.mark.parametrize("impl", [SIMT, TC])
@pytest.mark.parametrize("n,cin,cout,K,strided", [
    (700, 64, 64, 3, False),       # layer1 shape
    (1500, 64, 128, 3, True),      # strided conv
    (300, 128, 256, 3, False),
    (130, 256, 512, 3, False),     # two N tiles of 256
    (1000, 3, 64, 7, False),       # the k7 stem (SMALL mode)
    (129, 32, 64, 1, False),       # one row past a tile boundary
])
def test_conv_forward(cuda, impl, n, cin, cout, K, strided):
    c, out, nbr = _maps(n, K=K, strided=strided, extent=7 if K == 7 else 9)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((c.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((K ** 3, cin, cout)) * 0.05).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    ref = oo.conv(torch.from_numpy(x).double(), torch.from_numpy(w).double(), nbr, torch.from_numpy(b).double())
    got = _run_fwd(x, w, b, nbr, c.shape[0], out.shape[0], impl, cuda)
    util.assert_close(got, ref, tol=_tol(impl), what=f"conv fwd impl={impl}")
    if impl == SIMT:   # fp32 SIMT must be far tighter than the TF32 bar
        util.assert_close(got, ref, tol=2e-5, what="conv fwd simt fp32")
    else:
        with tf32_model():
            ref32 = oo.conv(torch.from_numpy(x), torch.from_numpy(w), nbr, torch.from_numpy(b))
        util.assert_close(got, ref32, tol=TF32_MODEL_TOL, what="conv fwd tcgen05 vs tf32 model")


@pytest.mark.parametrize("impl", [SIMT, TC])
def test_conv_use_mm_identity_map(cuda, impl):
    rng = np.random.default_rng(2)
    n, cin, cout = 777, 256, 64
    x = rng.standard_normal((n, cin)).astype(np.float32)
    w = (rng.standard_normal((1, cin, cout)) * 0.05).astype(np.float32)
    got = _run_fwd(x, w, None, None, n, n, impl, cuda)
    util.assert_close(got, torch.from_numpy(x).double() @ torch.from_numpy(w[0]).double(), tol=_tol(impl), what="use_mm")


@pytest.mark.parametrize("impl", [SIMT, TC])
@pytest.mark.parametrize("n,cin,cout,strided", [(900, 64, 64, False), (1200, 64, 128, True), (200, 128, 256, False)])
def test_conv_backward(cuda, impl, n, cin, cout, strided):
    """dgrad and wgrad through the autograd Function, against autograd of the oracle."""
    from dpcr_agb_b200.MinkowskiEngine.coordinate_manager import CoordinateManager
    c, out, nbr = _maps(n, strided=strided, seed=5)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((c.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * 0.05).astype(np.float32)
    b = rng.standard_normal((1, cout)).astype(np.float32)
    gy = rng.standard_normal((out.shape[0], cout)).astype(np.float32)

    xr = torch.from_numpy(x).double().requires_grad_()
    wr = torch.from_numpy(w).double().requires_grad_()
    br = torch.from_numpy(b).double().requires_grad_()
    oo.conv(xr, wr, nbr, br).backward(torch.from_numpy(gy).double())

    cm = CoordinateManager(D=3, device=cuda)
    key, _ = cm.insert(torch.from_numpy(c).to(cuda))
    out_key = cm.stride(key, 2) if strided else key
    assert np.array_equal(cm.coords(out_key).cpu().numpy(), out)
    km = cm.kernel_map(key, out_key, 3)
    old = Fn.CONV_IMPL
    Fn.CONV_IMPL = impl if impl == SIMT else 0     # auto: tcgen05 where covered (wgrad may still be SIMT)
    try:
        xg = torch.from_numpy(x).to(cuda).requires_grad_()
        wg = torch.from_numpy(w).to(cuda).requires_grad_()
        bg = torch.from_numpy(b).to(cuda).requires_grad_()
        y = Fn.ConvolutionFunction.apply(xg, wg, bg, km)
        y.backward(torch.from_numpy(gy).to(cuda))
    finally:
        Fn.CONV_IMPL = old
    util.assert_close(xg.grad, xr.grad, tol=_tol(impl), what="dgrad")
    util.assert_close(wg.grad, wr.grad, tol=_tol(impl), what="wgrad")
    util.assert_close(bg.grad, br.grad, what="bias grad")
    if impl == TC:
        x32 = torch.from_numpy(x).requires_grad_()
        w32 = torch.from_numpy(w).requires_grad_()
        with tf32_model():
            oo.conv(x32, w32, nbr, torch.from_numpy(b)).backward(torch.from_numpy(gy))
        util.assert_close(xg.grad, x32.grad, tol=TF32_MODEL_TOL, what="dgrad vs tf32 model")
        util.assert_close(wg.grad, w32.grad, tol=TF32_MODEL_TOL, what="wgrad vs tf32 model")


@pytest.mark.parametrize("impl", [SIMT, TC])
def test_stem_wgrad_small_cin(cuda, impl):
    c, out, nbr = _maps(1500, K=7, extent=7, seed=8)
    rng = np.random.default_rng(4)
    x = rng.standard_normal((c.shape[0], 3)).astype(np.float32)
    gy = rng.standard_normal((c.shape[0], 64)).astype(np.float32)
    xr = torch.from_numpy(x).double()
    wr = torch.zeros((343, 3, 64), dtype=torch.float64, requires_grad=True)
    oo.conv(xr, wr, nbr).backward(torch.from_numpy(gy).double())
    n = c.shape[0]
    got = Fn.wgrad(torch.from_numpy(x).to(cuda), torch.from_numpy(gy).to(cuda), torch.from_numpy(nbr).to(cuda),
                   n, n, 3, 64, 343, impl=impl)
    util.assert_close(got, wr.grad, tol=_tol(impl), what=f"stem wgrad impl={impl}")
    if impl == TC:
        w32 = torch.zeros((343, 3, 64), requires_grad=True)
        with tf32_model():
            oo.conv(torch.from_numpy(x), w32, nbr).backward(torch.from_numpy(gy))
        util.assert_close(got, w32.grad, tol=TF32_MODEL_TOL, what="stem wgrad vs tf32 model")


def test_conv_rejects_bad_arguments(cuda):
    x = torch.zeros((4, 8), device=cuda)
    w = torch.zeros((27, 8, 8), device=cuda)
    y = torch.zeros((4, 8), device=cuda)
    with pytest.raises(L.B2SError):   # nbr may be null only for k3 == 1
        L.call("b2s_conv_gather_gemm", x, w, None, None, 4, 4, None, 8, 8, 27, 0, y, None, 0, 1, None)
    with pytest.raises(L.B2SError):   # tcgen05 kernel does not cover c_out = 8
        L.call("b2s_conv_gather_gemm", x, w, None, None, 4, 4, None, 8, 8, 1, 0, y, None, 0, 2, None)


@pytest.mark.parametrize("n,cin,cout,K,strided", [
    (5000, 64, 64, 3, False),      # many row splits, atomics
    (3000, 64, 128, 3, True),
    (700, 128, 256, 3, False),
    (100, 512, 512, 3, False),     # single split: plain stores, 4 ci tiles x 2 co tiles
    (1000, 256, 64, 1, False),     # use_mm shape (identity map passed explicitly)
])
def test_wgrad_tcgen05(cuda, n, cin, cout, K, strided):
    c, out, nbr = _maps(n, K=K, strided=strided, seed=12)
    rng = np.random.default_rng(6)
    x = rng.standard_normal((c.shape[0], cin)).astype(np.float32)
    gy = rng.standard_normal((out.shape[0], cout)).astype(np.float32)
    xr = torch.from_numpy(x).double()
    wr = torch.zeros((K ** 3, cin, cout), dtype=torch.float64, requires_grad=True)
    oo.conv(xr, wr, nbr).backward(torch.from_numpy(gy).double())
    args = (torch.from_numpy(x).to(cuda), torch.from_numpy(gy).to(cuda), torch.from_numpy(nbr).to(cuda),
            c.shape[0], out.shape[0], cin, cout, K ** 3)
    got_tc = Fn.wgrad(*args, impl=TC)
    got_simt = Fn.wgrad(*args, impl=SIMT)
    util.assert_close(got_simt, wr.grad, tol=2e-5, what="wgrad simt")
    util.assert_close(got_tc, wr.grad, tol=_tol(TC), what="wgrad tcgen05")
    w32 = torch.zeros((K ** 3, cin, cout), requires_grad=True)
    with tf32_model():
        oo.conv(torch.from_numpy(x), w32, nbr).backward(torch.from_numpy(gy))
    util.assert_close(got_tc, w32.grad, tol=TF32_MODEL_TOL, what="wgrad tcgen05 vs tf32 model")


def test_wgrad_tcgen05_null_map(cuda):
    rng = np.random.default_rng(7)
    n, cin, cout = 900, 64, 256
    x = rng.standard_normal((n, cin)).astype(np.float32)
    gy = rng.standard_normal((n, cout)).astype(np.float32)
    got = Fn.wgrad(torch.from_numpy(x).to(cuda), torch.from_numpy(gy).to(cuda), None, n, n, cin, cout, 1, impl=TC)
    util.assert_close(got[0], torch.from_numpy(x).double().T @ torch.from_numpy(gy).double(), tol=_tol(TC),
                      what="wgrad use_mm")


def test_parity_plan_and_strided_dgrad(cuda):
    """b2s_parity_plan covers every fine row exactly once in tile-aligned parity classes, and the parity-plan dgrad
    equals the dense transposed-table dgrad (same products, different zero rows skipped)."""
    from dpcr_agb_b200.MinkowskiEngine.coordinate_manager import CoordinateManager
    rng = np.random.default_rng(21)
    c = util.random_coords(rng, 5000, nb=3, extent=14)
    cm = CoordinateManager(D=3, device=cuda)
    key, _ = cm.insert(torch.from_numpy(c).to(cuda))
    out_key = cm.stride(key, 2)
    for K, cin, cout in ((3, 64, 128), (1, 64, 128), (3, 128, 256)):
        km = cm.kernel_map(key, out_key, K)
        perm, bounds = km.parity_plan
        p = perm.cpu().numpy()
        b = bounds.cpu().numpy()
        live = p[p >= 0]
        assert np.array_equal(np.sort(live), np.arange(c.shape[0]))              # a permutation of the fine rows
        assert b[8] * 128 <= p.shape[0] and np.all(p[b[8] * 128:] == -1)
        cls = (c[:, 1] & 1) | ((c[:, 2] & 1) << 1) | ((c[:, 3] & 1) << 2)
        for k in range(8):
            rows = p[b[k] * 128:b[k + 1] * 128]
            assert np.all(cls[rows[rows >= 0]] == k)
        gy = Fn.round_tf32(torch.from_numpy(rng.standard_normal((km.n_out, cout)).astype(np.float32)).to(cuda))
        w = torch.from_numpy((rng.standard_normal((K ** 3, cin, cout)) * 0.05).astype(np.float32)).to(cuda)
        dense = Fn.gather_gemm(gy, w, None, km.inv, km.n_out, km.n_in, cout, cin, km.k3, 1, impl=TC, prerounded=True)
        sparse = Fn.dgrad_strided(gy, w, km, cout, cin)
        util.assert_close(sparse, dense, tol=2e-5, what=f"parity dgrad K={K}")   # fp32 summation order only


@pytest.mark.parametrize("n,cin,cout", [(700, 64, 64), (1300, 128, 128), (257, 64, 128), (900, 128, 64),
                                        (300, 128, 256), (400, 256, 512)])
def test_conv_m256_tiles(cuda, n, cin, cout):
    """The M = 256 variant of the tcgen05 kernel (two accumulator tiles per CTA sharing every weight stage; chosen
    automatically only for maps with >= 148 such tiles, forced here through the tuning knob): forward with bias and
    dgrad (transposed weights, reversed kernel index) against the oracle's tf32 model, incl. a partial last tile."""
    c, out, nbr = _maps(n, seed=7)
    rng = np.random.default_rng(8)
    x = rng.standard_normal((c.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * 0.05).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    gy = rng.standard_normal((out.shape[0], cout)).astype(np.float32)
    x32 = torch.from_numpy(x).requires_grad_()
    with tf32_model():
        ref = oo.conv(x32, torch.from_numpy(w), nbr, torch.from_numpy(b))
        ref.backward(torch.from_numpy(gy))
    ng = torch.from_numpy(nbr).to(cuda)
    wg = torch.from_numpy(w).to(cuda)
    L.set_tuning("tc_m256", 2)
    try:
        got = Fn.gather_gemm(torch.from_numpy(x).to(cuda), wg, torch.from_numpy(b).to(cuda), ng, c.shape[0],
                             out.shape[0], cin, cout, 27, 0, impl=TC)
        gx = Fn.gather_gemm(torch.from_numpy(gy).to(cuda), wg, None, ng, out.shape[0], c.shape[0], cout, cin, 27, 3,
                            impl=TC)
    finally:
        L.set_tuning("tc_m256", -1)
    util.assert_close(got, ref.detach(), tol=TF32_MODEL_TOL, what="M=256 forward vs tf32 model")
    util.assert_close(gx, x32.grad, tol=TF32_MODEL_TOL, what="M=256 dgrad vs tf32 model")


@pytest.mark.parametrize("n,cin,cout", [(700, 64, 64), (1300, 128, 64), (257, 64, 128), (900, 256, 64),
                                        (300, 128, 128), (520, 64, 192)])
def test_conv_operand_in_tensor_memory(cuda, n, cin, cout, opmode):
    """gather_gemm_ta_kernel (split-bf16 mode: gathered rows go registers -> tensor memory, weight image rows in the
    matching channel order; chosen automatically for 64-wide tiles of maps with >= 148 x 256 rows, forced here):
    forward with bias and dgrad against the oracle's precision model and against the shared-memory kernel, through a
    workspace image and through a prebuilt image (which holds both forms), incl. a partial last tile."""
    if opmode != "bf16x2":
        pytest.skip("split-bf16 operand mode only")
    c, out, nbr = _maps(n, seed=17)
    rng = np.random.default_rng(18)
    x = rng.standard_normal((c.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * 0.05).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    gy = rng.standard_normal((out.shape[0], cout)).astype(np.float32)
    x32 = torch.from_numpy(x).requires_grad_()
    with tf32_model():
        ref = oo.conv(x32, torch.from_numpy(w), nbr, torch.from_numpy(b))
        ref.backward(torch.from_numpy(gy))
    ng, wg, bg = torch.from_numpy(nbr).to(cuda), torch.from_numpy(w).to(cuda), torch.from_numpy(b).to(cuda)
    xg, gg = torch.from_numpy(x).to(cuda), torch.from_numpy(gy).to(cuda)
    n_in, n_out = c.shape[0], out.shape[0]

    def both(img_f=None, img_d=None):
        pre = img_f is not None
        xo, go = (Fn.round_tf32(xg), Fn.round_tf32(gg)) if pre else (xg, gg)
        y = Fn.gather_gemm(xo, wg, bg, ng, n_in, n_out, cin, cout, 27, 0, impl=TC, prerounded=pre, wimg=img_f)
        gx = Fn.gather_gemm(go, wg, None, ng, n_out, n_in, cout, cin, 27, 3, impl=TC, prerounded=pre, wimg=img_d)
        return y, gx

    L.set_tuning("tc_ta", 0)
    try:
        y0, gx0 = both()
        L.set_tuning("tc_ta", 3)
        y1, gx1 = both()
        imgs = []
        for layout, ci, co in ((0, cin, cout), (3, cout, cin)):
            nbytes = L.query("b2s_conv_weight_image_bytes", ci, co, 27)
            img = torch.empty(nbytes, dtype=torch.uint8, device=cuda)
            L.call("b2s_conv_weight_image", wg, ci, co, 27, layout, img, img.numel())
            imgs.append(img)
        y2, gx2 = both(*imgs)
        L.set_tuning("tc_ta", 0)
        y3, gx3 = both(*imgs)                        # the same prebuilt images serve the shared-memory kernel
    finally:
        L.set_tuning("tc_ta", -1)
    util.assert_close(y1, ref.detach(), tol=TF32_MODEL_TOL, what="forward vs precision model")
    util.assert_close(gx1, x32.grad, tol=TF32_MODEL_TOL, what="dgrad vs precision model")
    for what, a, r in (("forward", y1, y0), ("dgrad", gx1, gx0), ("forward, prebuilt image", y2, y0),
                       ("dgrad, prebuilt image", gx2, gx0), ("forward, prebuilt image, smem kernel", y3, y0),
                       ("dgrad, prebuilt image, smem kernel", gx3, gx0)):
        util.assert_close(a, r, tol=5e-6, what=what + " vs shared-memory kernel")   # same products, fp32 order only


@pytest.mark.parametrize("n,cin,cout,K", [
    (3000, 64, 64, 3),        # M = 128 kernel, 64-wide tiles
    (3000, 64, 128, 3),       # M = 256 kernel (two row tiles per CTA)
    (200, 128, 256, 3),       # few row tiles: split-K launch -> the statistics come from the fallback reduction
    (1300, 64, 256, 1),       # K = 1 through a table
])
def test_conv_epilogue_batch_norm_statistics(cuda, n, cin, cout, K):
    """``col_stats`` of b2s_conv_gather_gemm: column sums and sums of squares of the output (bias included, rows beyond
    the live count excluded), whichever way the launch produces them -- the input of b2s_bn_finalize."""
    c, out, nbr = _maps(n, K=K, seed=9)
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal((c.shape[0], cin)).astype(np.float32)).to(cuda)
    w = torch.from_numpy((rng.standard_normal((K ** 3, cin, cout)) * 0.05).astype(np.float32)).to(cuda)
    b = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).to(cuda)
    ng = torch.from_numpy(nbr).to(cuda)
    for live in (None, n - 77):
        elems = L.query("b2s_conv_col_stats_elems", n, cout)
        stats = torch.full((elems,), 123.0, dtype=torch.float32, device=cuda)
        n_dev = None if live is None else torch.tensor([live], dtype=torch.int32, device=cuda)
        y = Fn.gather_gemm(x, w, b, ng, n, n, cin, cout, K ** 3, 0, impl=TC, n_out_dev=n_dev, col_stats=stats)
        rows = n if live is None else live
        yd = y[:rows].double()
        rpt = int(stats[-4].item())                                  # header: out rows per partial row
        assert rpt in (128, 256)
        live_parts = -(-rows // rpt)
        parts = stats[:live_parts * 2 * cout].double().view(live_parts, 2, cout)
        s1, s2 = yd.sum(0), (yd * yd).sum(0)
        assert (parts[:, 0].sum(0) - s1).abs().max().item() <= 1e-5 * yd.abs().sum(0).max().item()
        assert (parts[:, 1].sum(0) - s2).abs().max().item() <= 1e-5 * s2.max().item()
        # ... and b2s_bn_finalize turns them into what b2s_bn_stats computes from y itself
        mean, invstd = torch.empty(cout, device=cuda), torch.empty(cout, device=cuda)
        mean2, invstd2 = torch.empty(cout, device=cuda), torch.empty(cout, device=cuda)
        rm, rv = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
        rm2, rv2 = torch.zeros(cout, device=cuda), torch.ones(cout, device=cuda)
        L.call("b2s_bn_finalize", stats, n, n_dev, cout, 1e-5, 0.1, rm, rv, mean, invstd)
        ws = torch.empty(2 * cout + 1, dtype=torch.float64, device=cuda)
        L.call("b2s_bn_stats", y, n, n_dev, cout, 1e-5, 0.1, rm2, rv2, ws, mean2, invstd2)
        util.assert_close(mean, mean2, tol=1e-6, what="fused mean")
        util.assert_close(invstd, invstd2, tol=1e-5, what="fused invstd")
        util.assert_close(rm, rm2, tol=1e-6, what="fused running mean")
        util.assert_close(rv, rv2, tol=1e-5, what="fused running var")
