"""GPU parity of the bandwidth-bound ops (a9-a16) against the oracle / torch fp64, forward and backward."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from dpcr_agb_b200.MinkowskiEngine.coordinate_manager import CoordinateManager
from oracle import coords as oc
from oracle import ops as oo
import b2s_testutil as util

pytestmark = pytest.mark.gpu


def _setup(cuda, n=4000, nb=3, seed=0):
    rng = np.random.default_rng(seed)
    c = util.random_coords(rng, n, nb=nb, extent=10)
    cm = CoordinateManager(D=3, device=cuda)
    key, _ = cm.insert(torch.from_numpy(c).to(cuda))
    return rng, c, cm, key


@pytest.mark.parametrize("ch", [64, 32, 128, 48, 3])
def test_maxpool_fwd_bwd(cuda, ch):
    """ch = 32 / 64 / 128: the warp-per-32-rows kernel (table entries staged coalesced, only existing neighbours
    visited); 48 and 3: the row-loop kernel (vector / scalar)."""
    rng, c, cm, key = _setup(cuda)
    out_key = cm.stride(key, 2)
    out_c = cm.coords(out_key).cpu().numpy()
    nbr = oc.kernel_map_table(c, out_c, 3, (1, 1, 1))
    x = rng.standard_normal((c.shape[0], ch)).astype(np.float32)
    x[:50] = 0.25                                   # ties -> lowest in-row
    gy = rng.standard_normal((out_c.shape[0], ch)).astype(np.float32)
    xr = torch.from_numpy(x).double().requires_grad_()
    yr = oo.max_pool(xr, nbr)
    yr.backward(torch.from_numpy(gy).double())
    xg = torch.from_numpy(x).to(cuda).requires_grad_()
    yg = Fn.MaxPoolFunction.apply(xg, cm.kernel_map(key, out_key, 3))
    yg.backward(torch.from_numpy(gy).to(cuda))
    assert torch.equal(yg.detach().cpu().double(), yr.detach())          # max of fp32 values is exact
    util.assert_close(xg.grad, xr.grad, tol=1e-6, what="maxpool bwd")


@pytest.mark.parametrize("avg", [False, True])
@pytest.mark.parametrize("ch", [64, 512, 48])
def test_global_pool_fwd_bwd(cuda, avg, ch):
    rng, c, cm, key = _setup(cuda, seed=1)
    x = rng.standard_normal((c.shape[0], ch)).astype(np.float32)
    g = rng.standard_normal((3, ch)).astype(np.float32)
    xr = torch.from_numpy(x).double().requires_grad_()
    yr = oo.global_pool(xr, c[:, 0], 3, "avg" if avg else "sum")
    yr.backward(torch.from_numpy(g).double())
    xg = torch.from_numpy(x).to(cuda).requires_grad_()
    yg = Fn.GlobalPoolFunction.apply(xg, cm.coords(key), 3, cm.inv_counts(key) if avg else None)
    yg.backward(torch.from_numpy(g).to(cuda))
    util.assert_close(yg, yr, tol=1e-5, what="global pool fwd")
    util.assert_close(xg.grad, xr.grad, tol=1e-6, what="global pool bwd")
    assert cm.rows_per_batch(key) == np.bincount(c[:, 0], minlength=3).tolist()


def test_broadcast_mul_fwd_bwd(cuda):
    rng, c, cm, key = _setup(cuda, seed=2)
    x = rng.standard_normal((c.shape[0], 128)).astype(np.float32)
    y = rng.standard_normal((3, 128)).astype(np.float32)
    g = rng.standard_normal(x.shape).astype(np.float32)
    xr, yr = torch.from_numpy(x).double().requires_grad_(), torch.from_numpy(y).double().requires_grad_()
    oo.broadcast_mul(xr, yr, c[:, 0]).backward(torch.from_numpy(g).double())
    xg, yg = torch.from_numpy(x).to(cuda).requires_grad_(), torch.from_numpy(y).to(cuda).requires_grad_()
    out = Fn.BroadcastMulFunction.apply(xg, yg, cm.coords(key), 3)
    out.backward(torch.from_numpy(g).to(cuda))
    util.assert_close(out, torch.from_numpy(x).double() * torch.from_numpy(y).double()[c[:, 0]], tol=1e-6, what="bmul")
    util.assert_close(xg.grad, xr.grad, tol=1e-6, what="bmul gx")
    util.assert_close(yg.grad, yr.grad, tol=1e-5, what="bmul gy")
    # per-plot scalar mask (drop path)
    m = torch.tensor([[1.0], [0.0], [1.0 / 0.99]])
    out2 = Fn.BroadcastMulFunction.apply(xg, m.to(cuda), cm.coords(key), 3)
    util.assert_close(out2, torch.from_numpy(x).double() * m.double()[c[:, 0]], tol=1e-6, what="mask mul")


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("training", [True, False])
def test_batch_norm_fwd_bwd(cuda, act, training):
    rng = np.random.default_rng(3)
    n, ch = 5000, 96
    x = (rng.standard_normal((n, ch)) * 2.0 + 5.0).astype(np.float32)      # non-zero mean: cancellation check
    w = rng.standard_normal(ch).astype(np.float32)
    b = rng.standard_normal(ch).astype(np.float32)
    g = rng.standard_normal((n, ch)).astype(np.float32)
    rm = rng.standard_normal(ch).astype(np.float32)
    rv = (rng.random(ch) + 0.5).astype(np.float32)

    xr = torch.from_numpy(x).double().requires_grad_()
    wr, br = torch.from_numpy(w).double().requires_grad_(), torch.from_numpy(b).double().requires_grad_()
    rmr, rvr = torch.from_numpy(rm).double(), torch.from_numpy(rv).double()
    yr = F.batch_norm(xr, rmr, rvr, wr, br, training, 0.1, 1e-5)
    if act:
        yr = F.gelu(yr)
    yr.backward(torch.from_numpy(g).double())

    xg = torch.from_numpy(x).to(cuda).requires_grad_()
    wg, bg = torch.from_numpy(w).to(cuda).requires_grad_(), torch.from_numpy(b).to(cuda).requires_grad_()
    rmg, rvg = torch.from_numpy(rm).to(cuda), torch.from_numpy(rv).to(cuda)
    yg = Fn.BatchNormFunction.apply(xg, wg, bg, rmg, rvg, training, 0.1, 1e-5, act)
    yg.backward(torch.from_numpy(g).to(cuda))
    util.assert_close(yg, yr, tol=1e-5, what="bn fwd")
    util.assert_close(xg.grad, xr.grad, tol=1e-4, what="bn gx")
    util.assert_close(wg.grad, wr.grad, tol=1e-4, what="bn gw")
    util.assert_close(bg.grad, br.grad, tol=1e-4, what="bn gb")
    util.assert_close(rmg, rmr, tol=1e-5, what="running mean")
    util.assert_close(rvg, rvr, tol=1e-5, what="running var")


@pytest.mark.parametrize("n,ch,training", [(5000, 64, False), (3001, 256, False), (777, 2048, False), (5000, 96, False),
                                           (4000, 128, True)])
def test_bn_backward_column_sums_of_gx(cuda, n, ch, training):
    """b2s_bn_bwd_apply's gx_colsum output (the bias gradient of the convolution in front of the norm) equals the
    column sums of gx: fused form (power-of-two channel counts), the internal fallback (96 channels), and the
    training-mode case where the sums cancel to rounding noise.  ConvolutionFunction picks the attribute up."""
    rng = np.random.default_rng(11)
    x = torch.from_numpy((rng.standard_normal((n, ch)) + 1.0).astype(np.float32)).to(cuda).requires_grad_()
    w = torch.from_numpy(rng.standard_normal(ch).astype(np.float32)).to(cuda).requires_grad_()
    b = torch.from_numpy(rng.standard_normal(ch).astype(np.float32)).to(cuda).requires_grad_()
    rm, rv = torch.zeros(ch, device=cuda), torch.ones(ch, device=cuda)
    y = Fn.BatchNormFunction.apply(x, w, b, rm, rv, training, 0.1, 1e-5, 1)
    g = torch.from_numpy(rng.standard_normal((n, ch)).astype(np.float32)).to(cuda)
    seen = {}
    x.register_hook(lambda gx: seen.update(cs=getattr(gx, "_b2s_colsum", None), gx=gx))
    y.backward(g)
    cs, gx = seen["cs"], seen["gx"]
    assert cs is not None and cs[1] == gx._version and cs[0].dim() == 2 and cs[0].shape[1] == ch
    ref = gx.double().sum(0)
    scale = gx.double().abs().sum(0).max().item()               # the sums cancel in training mode: compare on this scale
    assert (cs[0].double().sum(0) - ref).abs().max().item() <= 2e-6 * scale


def test_gelu_fwd_bwd(cuda):
    x = torch.linspace(-6, 6, 10001)
    g = torch.randn(10001, generator=torch.Generator().manual_seed(0))
    xr = x.double().requires_grad_()
    F.gelu(xr).backward(g.double())
    xg = x.to(cuda).requires_grad_()
    yg = Fn.GELUFunction.apply(xg)
    yg.backward(g.to(cuda))
    util.assert_close(yg, F.gelu(x.double()), tol=1e-6, what="gelu")
    util.assert_close(xg.grad, xr.grad, tol=1e-6, what="gelu grad")
