"""GPU parity of the x-LINE path of the k7 stem (SURVEY.md 8 rows a4, a6, a8 for ME/SENet.py:49-52): the line table
``b2s_kernel_map_lines`` bit-exact against the oracle, and the line-driven forward / weight-gradient kernels
(``b2s_conv_lines_fwd`` / ``_wgrad``, split-bf16 operands) against the fp64 oracle, against the oracle's precision
model of that mode, and against the table-driven kernels they replace.  Everything goes through the C ABI."""
import numpy as np
import pytest
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import lib as L
from dpcr_agb_b200.MinkowskiEngine import coordinate_manager as CM
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from dpcr_agb_b200.quantize import GridSampling3D
from oracle import coords as oc
from oracle import ops as oo
import b2s_testutil as util

pytestmark = pytest.mark.gpu


def _voxels(cuda, num_plots, n_points, size, cfg=17, bounds=None):
    batch = util.make_points(num_plots, n_points, cfg=cfg)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) for k, v in batch.items()}
    vox = GridSampling3D(size)(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=num_plots,
                               bounds=bounds)
    return vox


def _stem_map(cuda, vox, K):
    cm = CM.CoordinateManager(D=3, device=cuda)
    key, _ = cm.insert(vox["coords"], dense_index=vox["index"])
    return cm, cm.kernel_map(key, key, K)


@pytest.mark.parametrize("num_plots,n_points,size,K", [
    (3, 3000, 0.02, 7),               # the stem shape
    (2, 16000, 0.0125, 7),            # BASELINE plot size
    (2, 2000, 0.05, 3),
    (1, 1500, 0.04, (8, 3, 2)),       # widest line, even kernel sizes (offsets 0 .. K-1)
    (2, 1200, 0.05, (1, 5, 5)),
])
def test_line_table_bit_exact(cuda, num_plots, n_points, size, K):
    vox = _voxels(cuda, num_plots, n_points, size)
    cm, km = _stem_map(cuda, vox, K)
    assert km.lines_ok and km._nbr is None, "the x-line map must not build the [K^3, N] table up front"
    c = vox["coords"].cpu().numpy()
    ref_tab = oc.kernel_map_table(c, c, km.kernel_size, (1, 1, 1))
    ref = oc.kernel_map_lines(ref_tab, km.kernel_size)
    got = km.lines.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, ref)
    assert np.array_equal(oc.lines_to_table(got, km.kernel_size), km.nbr.cpu().numpy())    # lazily built table agrees
    assert km.num_pairs() == int((ref_tab >= 0).sum()) > c.shape[0]


def test_line_table_tight_bounds_and_static_rows(cuda):
    """Voxels on the faces of the quantiser's box (windows clipped at x = 0 / dim - 1, lines outside in y / z) and the
    static form: rows allocated at a capacity, live count on the device, entries beyond it untouched."""
    vox = _voxels(cuda, 2, 2500, 0.03)           # measured bounds: the box is exactly the occupied range
    cm, km = _stem_map(cuda, vox, 7)
    c = vox["coords"].cpu().numpy()
    ref = oc.kernel_map_lines(oc.kernel_map_table(c, c, 7, (1, 1, 1)), 7)
    assert np.array_equal(km.lines.cpu().numpy().view(np.uint32), ref)
    n = c.shape[0]
    cap = n + 300
    ws, lo, dims, num_plots = vox["index"]
    coords_cap = torch.zeros((cap, 4), dtype=torch.int32, device=cuda)
    coords_cap[:n] = vox["coords"]
    lines = torch.full((49, cap), -7, dtype=torch.int32, device=cuda)
    n_dev = torch.tensor([n], dtype=torch.int32, device=cuda)
    L.call("b2s_kernel_map_lines", coords_cap, cap, n_dev, ws, num_plots, L.host_i32(*lo), L.host_i32(*dims),
           L.host_i32(7, 7, 7), L.host_i32(1, 1, 1), lines)
    got = lines.cpu().numpy()
    assert np.array_equal(got[:, :n].view(np.uint32), ref) and np.all(got[:, n:] == -7)


def _conv_inputs(vox, c_out, K, seed=3):
    rng = np.random.default_rng(seed)
    n = vox["coords"].shape[0]
    k3 = K ** 3
    w = (rng.standard_normal((k3, 3, c_out)) * 0.05).astype(np.float32)
    b = rng.standard_normal(c_out).astype(np.float32)
    gy = rng.standard_normal((n, c_out)).astype(np.float32)
    return w, b, gy


@pytest.mark.parametrize("num_plots,n_points,size,c_out,K", [
    (3, 3000, 0.02, 64, 7),           # the stem
    (2, 700, 0.05, 128, 7),           # two output-channel tiles, fewer rows than one 256-row tile per plot
    (2, 2500, 0.03, 64, 3),           # 9 lines: two line groups of 5 + 4 in the weight gradient
])
def test_line_conv_forward_and_wgrad(cuda, num_plots, n_points, size, c_out, K):
    L.set_tuning("precise", 1)
    try:
        vox = _voxels(cuda, num_plots, n_points, size)
        cm, km = _stem_map(cuda, vox, K)
        x = vox["tensors"][0]
        assert x.shape[1] == 3 and Fn.lines_path(km, 3, c_out)
        w, b, gy = _conv_inputs(vox, c_out, K)
        wg, bg, gyg = (torch.from_numpy(a).to(cuda) for a in (w, b, gy))
        c = vox["coords"].cpu().numpy()
        nbr = oc.kernel_map_table(c, c, K, (1, 1, 1))
        xd = x.cpu().double().requires_grad_()
        wd = torch.from_numpy(w).double().requires_grad_()
        ref = oo.conv(xd, wd, nbr, torch.from_numpy(b).double())
        ref.backward(torch.from_numpy(gy).double())
        # --- forward
        y = Fn.lines_fwd(x, wg, bg, km, 3, c_out)
        util.assert_close(y, ref, tol=2e-6, what="line conv forward vs fp64 oracle")
        y_tab = Fn.gather_gemm(x, wg, bg, km.nbr, km.n_in, km.n_out, 3, c_out, km.k3, 0, impl=2)
        util.assert_close(y, y_tab, tol=2e-6, what="line conv forward vs table-driven kernel")
        # --- weight gradient (gy in operand form, as the autograd Function passes it)
        gw = Fn.lines_wgrad(x, Fn.round_tf32(gyg), km, 3, c_out)
        util.assert_close(gw, wd.grad, tol=5e-5, what="line conv wgrad vs fp64 oracle")
        old = oo.CONV_PRECISION
        oo.CONV_PRECISION = "bf16x2"
        try:
            x32 = x.cpu().requires_grad_()
            w32 = torch.from_numpy(w).requires_grad_()
            oo.conv(x32, w32, nbr, torch.from_numpy(b)).backward(torch.from_numpy(gy))
        finally:
            oo.CONV_PRECISION = old
        util.assert_close(gw, w32.grad, tol=2e-5, what="line conv wgrad vs the oracle's split-bf16 model")
        gw_tab = Fn.wgrad(x, Fn.round_tf32(gyg), km.nbr, km.n_in, km.n_out, 3, c_out, km.k3, impl=2, prerounded=True)
        util.assert_close(gw, gw_tab, tol=2e-6, what="line conv wgrad vs table-driven kernel")
    finally:
        L.set_tuning("precise", -1)


def test_line_conv_through_the_module_and_static_rows(cuda):
    """MinkowskiConvolution(3, 64, kernel_size=7) on a quantised SparseTensor takes the line path (no [343, N] table
    is built), its backward matches the oracle, and the static form (capacity rows + device count) gives the same
    rows as the exact form."""
    L.set_tuning("precise", 1)
    try:
        vox = _voxels(cuda, 2, 4000, 0.02, bounds=((0, 0, 0), (50, 50, 70)))
        torch.manual_seed(0)
        conv = ME.MinkowskiConvolution(3, 64, kernel_size=7, stride=1, bias=True, dimension=3).to(cuda)
        st = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], dense_index=vox["index"])
        y = conv(st)
        km = st.coordinate_manager.kernel_map(st.coordinate_map_key, st.coordinate_map_key, 7)
        assert km._lines is not None and km._nbr is None
        gy = torch.randn_like(y.F)
        y.F.backward(gy)
        assert km._nbr is None, "backward of the stem must not build the [343, N] table either"
        c = vox["coords"].cpu().numpy()
        nbr = oc.kernel_map_table(c, c, 7, (1, 1, 1))
        wd = conv.kernel.detach().cpu().double().requires_grad_()
        bd = conv.bias.detach().cpu().double().requires_grad_()
        ref = oo.conv(vox["tensors"][0].cpu().double(), wd, nbr, bd)
        ref.backward(gy.cpu().double())
        util.assert_close(y.F, ref, tol=2e-6, what="module forward")
        util.assert_close(conv.kernel.grad, wd.grad, tol=5e-5, what="module kernel gradient")
        util.assert_close(conv.bias.grad, bd.grad, tol=1e-5, what="module bias gradient")
        # --- static rows: same kernels at a capacity with the live count on the device
        n = c.shape[0]
        cap = n + 777
        xs = torch.zeros((cap, 3), device=cuda)
        xs[:n] = vox["tensors"][0]
        cs = torch.zeros((cap, 4), dtype=torch.int32, device=cuda)
        cs[:n] = vox["coords"]
        cmS = CM.CoordinateManager(D=3, device=cuda, capacities={1: cap}, num_batches=2)
        keyS = cmS.insert_static(cs, vox["num_rows"], dense_index=vox["index"])
        kmS = cmS.kernel_map(keyS, keyS, 7)
        assert kmS.lines_ok
        yS = Fn.lines_fwd(xs, conv.kernel.detach(), conv.bias.detach().view(-1), kmS, 3, 64)
        assert torch.equal(yS[:n], y.F.detach())
        gyS = torch.zeros((cap, 64), device=cuda)
        gyS[:n] = gy
        gwS = Fn.lines_wgrad(xs, Fn.round_tf32(gyS, vox["num_rows"]), kmS, 3, 64)
        util.assert_close(gwS, conv.kernel.grad, tol=2e-6, what="static-row wgrad (atomics: summation order differs)")
    finally:
        L.set_tuning("precise", -1)


def test_line_path_is_off_in_tf32_mode_and_for_wide_inputs(cuda):
    vox = _voxels(cuda, 1, 1500, 0.04)
    cm, km = _stem_map(cuda, vox, 7)
    L.set_tuning("precise", 0)
    try:
        assert not Fn.lines_path(km, 3, 64)          # TF32 operand mode keeps the table-driven kernels
    finally:
        L.set_tuning("precise", -1)
    assert not Fn.lines_path(km, 64, 64) and not Fn.lines_path(km, 3, 48)
    with pytest.raises(L.B2SError):
        ws = torch.empty(1 << 20, dtype=torch.uint8, device=cuda)
        L.call("b2s_conv_lines_fwd", vox["tensors"][0], torch.zeros(343, 3, 48, device=cuda), None, km.lines, km.n_in,
               km.n_out, None, 3, 48, L.host_i32(7, 7, 7), torch.empty(km.n_out, 48, device=cuda), ws, 1 << 20, None)


def test_line_conv_epilogue_batch_norm_statistics(cuda):
    """col_stats of b2s_conv_lines_fwd == column sums / sums of squares of its output over the live rows."""
    L.set_tuning("precise", 1)
    try:
        vox = _voxels(cuda, 2, 3000, 0.02)
        cm, km = _stem_map(cuda, vox, 7)
        w, b, _ = _conv_inputs(vox, 64, 7)
        n = km.n_out
        stats = torch.full((L.query("b2s_conv_col_stats_elems", n, 64),), -5.0, dtype=torch.float32, device=cuda)
        y = Fn.lines_fwd(vox["tensors"][0], torch.from_numpy(w).to(cuda), torch.from_numpy(b).to(cuda), km, 3, 64,
                         col_stats=stats)
        yd = y.double()
        assert int(stats[-4].item()) == 256
        parts = stats[:-(-n // 256) * 128].double().view(-1, 2, 64)
        assert (parts[:, 0].sum(0) - yd.sum(0)).abs().max().item() <= 1e-5 * yd.abs().sum(0).max().item()
        assert (parts[:, 1].sum(0) - (yd * yd).sum(0)).abs().max().item() <= 1e-5 * (yd * yd).sum(0).max().item()
    finally:
        L.set_tuning("precise", -1)
