"""GPU parity of the SURVEY.md 8(f) rank-3 ops against the oracle: non-generative transposed convolution (the MinkUNet
decoder, networks.py:155-176), local average / sum pooling (networks.py:29), pooling transpose, union-map addition --
forward and backward, through the same C ABI kernels as the MSENet path."""
import numpy as np
import pytest
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from oracle import me_cpu
import b2s_testutil as util

pytestmark = pytest.mark.gpu


def _inputs(n, c, seed=0, nb=2, extent=10):
    rng = np.random.default_rng(seed)
    coords = util.random_coords(rng, n, nb=nb, extent=extent)
    feats = rng.standard_normal((coords.shape[0], c)).astype(np.float32)
    return coords, feats


def _pair(coords, feats, cuda):
    xr = me_cpu.SparseTensor(torch.from_numpy(feats).clone().requires_grad_(), coordinates=torch.from_numpy(coords))
    xg = ME.SparseTensor(torch.from_numpy(feats).to(cuda).requires_grad_(), coordinates=torch.from_numpy(coords).to(cuda))
    return xr, xg


def _copy_params(dst, src, cuda):
    dst.load_state_dict(src.state_dict())
    return dst.to(cuda)


@pytest.mark.parametrize("K,cin,cmid,cout", [(2, 32, 64, 64), (3, 64, 64, 128), (2, 16, 32, 24)])
def test_conv_transpose_onto_the_encoder_map(cuda, K, cin, cmid, cout):
    """conv(K, stride 2) down, transposed conv(K, stride 2) back up onto the encoder map, as MinkUNet does."""
    coords, feats = _inputs(1500, cin, seed=K)
    torch.manual_seed(K)
    down_r = me_cpu.MinkowskiConvolution(cin, cmid, kernel_size=K, stride=2, dimension=3)
    up_r = me_cpu.MinkowskiConvolutionTranspose(cmid, cout, kernel_size=K, stride=2, bias=True, dimension=3)
    down_g = _copy_params(ME.MinkowskiConvolution(cin, cmid, kernel_size=K, stride=2, dimension=3), down_r, cuda)
    up_g = _copy_params(ME.MinkowskiConvolutionTranspose(cmid, cout, kernel_size=K, stride=2, bias=True, dimension=3),
                        up_r, cuda)
    xr, xg = _pair(coords, feats, cuda)
    yr, yg = up_r(down_r(xr)), up_g(down_g(xg))
    assert yg.coordinate_map_key == xg.coordinate_map_key and yg.F.shape == (coords.shape[0], cout)
    assert np.array_equal(yg.C.cpu().numpy(), yr.C)
    util.assert_close(yg.F, yr.F, tol=5e-5, what="transposed convolution forward")
    g = torch.randn(yr.F.shape, generator=torch.Generator().manual_seed(1))
    yr.F.backward(g)
    yg.F.backward(g.to(cuda))
    util.assert_close(xg.F.grad, xr.F.grad, tol=5e-5, what="grad through conv + transposed conv")
    util.assert_close(up_g.kernel.grad, up_r.kernel.grad, tol=5e-5, what="transposed kernel grad")
    util.assert_close(up_g.bias.grad, up_r.bias.grad, tol=5e-5, what="transposed bias grad")
    util.assert_close(down_g.kernel.grad, down_r.kernel.grad, tol=5e-5, what="down kernel grad")


def test_conv_transpose_needs_an_existing_map(cuda):
    coords, feats = _inputs(200, 8)
    _, xg = _pair(coords, feats, cuda)
    up = ME.MinkowskiConvolutionTranspose(8, 8, kernel_size=2, stride=2, dimension=3).to(cuda)
    with pytest.raises((NotImplementedError, ValueError)):
        up(xg)                       # tensor stride 1 / 2: nothing finer to land on


@pytest.mark.parametrize("cls", ["MinkowskiAvgPooling", "MinkowskiSumPooling"])
@pytest.mark.parametrize("K,stride,c", [(2, 2, 64), (3, 2, 32), (3, 1, 20)])
def test_local_pooling(cuda, cls, K, stride, c):
    coords, feats = _inputs(1200, c, seed=K + stride)
    xr, xg = _pair(coords, feats, cuda)
    yr = getattr(me_cpu, cls)(kernel_size=K, stride=stride, dimension=3)(xr)
    yg = getattr(ME, cls)(kernel_size=K, stride=stride, dimension=3)(xg)
    assert np.array_equal(yg.C.cpu().numpy(), yr.C)
    util.assert_close(yg.F, yr.F, tol=2e-6, what=f"{cls} forward")
    g = torch.randn(yr.F.shape, generator=torch.Generator().manual_seed(2))
    yr.F.backward(g)
    yg.F.backward(g.to(cuda))
    util.assert_close(xg.F.grad, xr.F.grad, tol=2e-6, what=f"{cls} backward")


def test_pooling_transpose_is_the_adjoint_of_sum_pooling(cuda):
    """<unpool(a), b> == <a, sumpool(b)> on the same pair of maps."""
    coords, feats = _inputs(900, 16, seed=4)
    _, xg = _pair(coords, feats, cuda)
    pooled = ME.MinkowskiSumPooling(kernel_size=2, stride=2, dimension=3)(xg)
    a = torch.randn_like(pooled.F)
    up = ME.MinkowskiPoolingTranspose(kernel_size=2, stride=2, dimension=3)(
        ME.SparseTensor(a, coordinate_map_key=pooled.coordinate_map_key, coordinate_manager=pooled.coordinate_manager))
    assert up.coordinate_map_key == xg.coordinate_map_key
    lhs = (up.F.double() * xg.F.detach().double()).sum().item()
    rhs = (a.double() * pooled.F.detach().double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0)


def test_union_map_addition(cuda):
    """``a + b`` / ``a - b`` for tensors on different coordinate maps of one manager (``SparseTensor.__add__``)."""
    rng = np.random.default_rng(9)
    ca = util.random_coords(rng, 400, nb=2, extent=6)
    cb = util.random_coords(rng, 500, nb=2, extent=6)
    fa = rng.standard_normal((400, 8)).astype(np.float32)
    fb = rng.standard_normal((500, 8)).astype(np.float32)
    ar = me_cpu.SparseTensor(torch.from_numpy(fa), coordinates=torch.from_numpy(ca))
    kb = ar.coordinate_manager.insert(cb, (1, 1, 1), tag="b")
    br = me_cpu.SparseTensor(torch.from_numpy(fb), coordinate_map_key=kb, coordinate_manager=ar.coordinate_manager)
    ag = ME.SparseTensor(torch.from_numpy(fa).to(cuda).requires_grad_(), coordinates=torch.from_numpy(ca).to(cuda))
    kbg, _ = ag.coordinate_manager.insert(torch.from_numpy(cb).to(cuda), (1, 1, 1), tag="b")
    bg = ME.SparseTensor(torch.from_numpy(fb).to(cuda).requires_grad_(), coordinate_map_key=kbg,
                         coordinate_manager=ag.coordinate_manager)
    for sign, op in ((1, lambda x, y: x + y), (-1, lambda x, y: x - y)):
        sr, sg = (ar + br) if sign == 1 else ar._union_add(br, -1.0), op(ag, bg)
        assert np.array_equal(sg.C.cpu().numpy(), sr.C)
        assert torch.equal(sg.F.detach().cpu(), sr.F)
    (ag + bg).F.sum().backward()
    assert torch.equal(ag.F.grad, torch.ones_like(ag.F)) and torch.equal(bg.F.grad, torch.ones_like(bg.F))


def test_minkunet_shaped_network(cuda):
    """A two-level MinkUNet-shaped network (networks.py:130-245 pattern: k2s2 down convolutions, k2s2 transposed
    convolutions back onto the encoder maps, ME.cat with the skip tensors, a residual block per level) forward +
    backward against the same modules over the oracle."""
    def build(M):
        torch.manual_seed(3)
        return torch.nn.ModuleDict({
            "c0": M.MinkowskiConvolution(3, 32, kernel_size=5, dimension=3), "b0": M.MinkowskiBatchNorm(32),
            "d1": M.MinkowskiConvolution(32, 64, kernel_size=2, stride=2, dimension=3), "bd1": M.MinkowskiBatchNorm(64),
            "d2": M.MinkowskiConvolution(64, 128, kernel_size=2, stride=2, dimension=3), "bd2": M.MinkowskiBatchNorm(128),
            "u2": M.MinkowskiConvolutionTranspose(128, 64, kernel_size=2, stride=2, dimension=3),
            "bu2": M.MinkowskiBatchNorm(64),
            "m2": M.MinkowskiConvolution(128, 64, kernel_size=3, dimension=3),
            "u1": M.MinkowskiConvolutionTranspose(64, 32, kernel_size=2, stride=2, dimension=3),
            "bu1": M.MinkowskiBatchNorm(32),
            "final": M.MinkowskiConvolution(64, 5, kernel_size=1, bias=True, dimension=3), "relu": M.MinkowskiReLU()})

    def run(M, net, x):
        r = net["relu"]
        p1 = r(net["b0"](net["c0"](x)))
        p2 = r(net["bd1"](net["d1"](p1)))
        p4 = r(net["bd2"](net["d2"](p2)))
        o = r(net["bu2"](net["u2"](p4)))
        o = r(net["m2"](M.cat(o, p2)))
        o = r(net["bu1"](net["u1"](o)))
        return net["final"](M.cat(o, p1))

    coords, feats = _inputs(1800, 3, seed=6, extent=9)
    net_r = build(me_cpu)
    net_g = build(ME)
    net_g.load_state_dict(net_r.state_dict())
    net_g = net_g.to(cuda)
    xr, xg = _pair(coords, feats, cuda)
    yr, yg = run(me_cpu, net_r, xr), run(ME, net_g, xg)
    assert yg.coordinate_map_key == xg.coordinate_map_key
    util.assert_close(yg.F, yr.F, tol=1e-3, what="MinkUNet-shaped forward")
    g = torch.randn(yr.F.shape, generator=torch.Generator().manual_seed(5))
    yr.F.backward(g)
    yg.F.backward(g.to(cuda))
    gmax = max(p.grad.abs().max().item() for p in net_r.parameters())
    for (n1, p1), (_, p2) in zip(net_g.named_parameters(), net_r.named_parameters()):
        err = (p1.grad.cpu() - p2.grad).abs().max().item()
        assert err <= 1e-3 * p2.grad.abs().max().item() + 1e-5 * gmax, f"{n1}: {err:.3e}"
