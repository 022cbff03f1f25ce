"""CPU tests of the oracle itself: hand-derived small cases, the golden fixtures, and hypothesis
properties of the integer stages (SURVEY.md section 4 test plan (i) and (iii))."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import coords as oc
from oracle import me_cpu
from oracle import ops as oo
import b2s_testutil as util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_quantize_matches_torch_expression_fixture():
    """grid_transform.py:116 evaluated by torch CPU (fixture made by tests/golden/make_golden.py)."""
    d = np.load(os.path.join(GOLD, "gridsampling_ref.npz"))
    got = oc.quantize_points(d["pos"], float(d["size"]))
    assert np.array_equal(got, d["coords"])
    # and live against torch, including half-way cases in both directions
    pos = torch.tensor([[0.00625, 0.01875, 0.03125], [0.04375, -0.00625, 1.0]], dtype=torch.float32)
    assert np.array_equal(oc.quantize_points(pos.numpy(), 0.0125), torch.round(pos / 0.0125).numpy())


def test_quantize_plot_hand_case():
    # size 1: points 0,1 share voxel (0,0,0); 2 alone in (1,0,0); 3,4 share (0,0,1)
    pos = np.array([[0.1, 0.2, 0.0], [0.4, -0.3, 0.2], [1.2, 0.0, 0.1], [0.0, 0.1, 0.9], [-0.2, 0.3, 1.1]], np.float32)
    feats = np.arange(5, dtype=np.float32)[:, None]
    coords, f, p, src = oc.quantize_plot(pos, feats, 1.0, None)
    # sorted by (z, y, x); representative = last point of the voxel in the (identity) shuffled order
    assert coords.tolist() == [[0, 0, 0], [1, 0, 0], [0, 0, 1]]
    assert src.tolist() == [1, 2, 4] and f[:, 0].tolist() == [1.0, 2.0, 4.0]
    # with a permutation the representative is the last in SHUFFLED order: perm = [4,3,2,1,0] -> 0 wins over 1
    coords2, f2, _, src2 = oc.quantize_plot(pos, feats, 1.0, np.array([4, 3, 2, 1, 0]))
    assert coords2.tolist() == coords.tolist() and src2.tolist() == [0, 2, 3]


def test_stride_map_hand_case():
    c = np.array([[0, 0, 0, 0], [0, 1, 0, 0], [0, 2, 0, 0], [0, -1, 0, 0], [1, 3, 3, -3], [0, 3, 0, 0]], np.int32)
    out, inv = oc.stride_map(c, (2, 2, 2))
    # floor toward -inf: -1 -> -2, -3 -> -4; first-occurrence order
    assert out.tolist() == [[0, 0, 0, 0], [0, 2, 0, 0], [0, -2, 0, 0], [1, 2, 2, -4]]
    assert inv.tolist() == [0, 0, 1, 2, 3, 1]


def test_kernel_map_and_conv_hand_case():
    # three voxels in a row along x (batch 0) and one isolated voxel in batch 1
    c = np.array([[0, 0, 0, 0], [0, 1, 0, 0], [0, 2, 0, 0], [1, 1, 0, 0]], np.int32)
    nbr = oc.kernel_map_table(c, c, 3, (1, 1, 1))
    assert nbr.shape == (27, 4)
    centre, left, right = 13, 12, 14                       # k = ix + 3 iy + 9 iz with (dx,dy,dz) = (ix-1, ...)
    assert nbr[centre].tolist() == [0, 1, 2, 3]
    assert nbr[left].tolist() == [-1, 0, 1, -1]            # in = out + (-1,0,0)
    assert nbr[right].tolist() == [1, 2, -1, -1]
    assert (nbr >= 0).sum() == 4 + 2 + 2
    i, o, off = oc.table_to_pairs(nbr)
    assert off[-1] == 8 and off[left] == 0 and off[centre] == 2 and off[right] == 6
    assert i[:2].tolist() == [0, 1] and o[:2].tolist() == [1, 2]
    # conv with W[k] = k * I : out[o] = sum_k k * x[nbr[k,o]]
    x = torch.tensor([[1.0], [10.0], [100.0], [7.0]])
    w = torch.arange(27, dtype=torch.float32).view(27, 1, 1)
    y = oo.conv(x, w, nbr)
    assert y[:, 0].tolist() == [13 * 1 + 14 * 10, 12 * 1 + 13 * 10 + 14 * 100, 12 * 10 + 13 * 100, 13 * 7]
    # strided map: out coords floor to multiples of 2; the k1 s2 "downsample" map has pairs only where the out
    # coordinate itself exists in the in map
    out, _ = oc.stride_map(c, (2, 2, 2))
    k1 = oc.kernel_map_table(c, out, 1, (1, 1, 1))
    assert out.tolist() == [[0, 0, 0, 0], [0, 2, 0, 0], [1, 0, 0, 0]] and k1.tolist() == [[0, 2, -1]]


def test_max_pool_ties_and_gradient():
    nbr = np.array([[0, 2], [1, -1]], np.int32)            # out0 <- {0,1}, out1 <- {2}
    x = torch.tensor([[1.0, 5.0], [1.0, 7.0], [3.0, 4.0]], requires_grad=True)
    y = oo.max_pool(x, nbr)
    assert y.tolist() == [[1.0, 7.0], [3.0, 4.0]]
    y.sum().backward()
    assert x.grad.tolist() == [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]   # tie in column 0 -> lowest in-row


def test_msenet14_oracle_fixture():
    """The committed golden input/output pair still reproduces (guards the oracle against drift)."""
    from dpcr_agb_b200 import msenet
    d = np.load(os.path.join(GOLD, "msenet14_oracle.npz"))
    pos_l = [d["pos"][d["batch"] == b] for b in range(2)]
    feat_l = [d["feats"][d["batch"] == b] for b in range(2)]
    perm_l = [d["perm"][:1200], d["perm"][1200:] - 1200]
    c, f, _, s, _ = oc.quantize_batch(pos_l, feat_l, 0.05, perm_l)
    assert np.array_equal(c, d["coords"]) and np.array_equal(s, d["src"]) and np.array_equal(f, d["vox_feats"])
    cur = c
    for ts in (2, 4, 8, 16):
        cur, _ = oc.stride_map(cur, (ts,) * 3)
        assert np.array_equal(cur, d[f"ts{ts}"])
    assert np.array_equal((oc.kernel_map_table(c, c, 7, (1, 1, 1)) >= 0).sum(1), d["stem_pairs"])
    assert np.array_equal(oc.kernel_map_table(c, d["ts2"], 3, (1, 1, 1)), d["pool_nbr"])
    torch.manual_seed(0)
    net = msenet.MSENet(me_cpu, "SENet14", drop_path=0.0).eval()
    with torch.no_grad():
        y = net(me_cpu.SparseTensor(torch.from_numpy(f), coordinates=torch.from_numpy(c))).numpy()
    assert np.allclose(y, d["output"], rtol=1e-4, atol=1e-5)


coord_lists = st.lists(st.tuples(st.integers(0, 2), st.integers(-9, 9), st.integers(-9, 9), st.integers(-9, 9)),
                       min_size=1, max_size=60, unique=True)


@settings(max_examples=60, deadline=None)
@given(coord_lists, st.sampled_from([1, 2, 3, 5]))
def test_kernel_map_properties(rows, K):
    c = np.asarray(rows, np.int32)
    nbr = oc.kernel_map_table(c, c, K, (1, 1, 1))
    offs = oc.kernel_offsets(K, (1, 1, 1))
    k3 = K ** 3
    for k in range(k3):
        o = np.nonzero(nbr[k] >= 0)[0]
        # definition: in = out + delta, same batch
        assert np.array_equal(c[nbr[k, o]][:, 1:], c[o][:, 1:] + offs[k])
        assert np.array_equal(c[nbr[k, o]][:, 0], c[o][:, 0])
        if K % 2 == 1:   # point symmetry of stride-1 odd kernels (used by dgrad)
            assert np.array_equal(nbr[k3 - 1 - k][nbr[k, o]], o)
    if K % 2 == 1:
        assert np.array_equal(nbr[k3 // 2], np.arange(c.shape[0]))
    # brute-force pair count
    keys = {tuple(r) for r in rows}
    want = sum((r[0], r[1] + d[0], r[2] + d[1], r[3] + d[2]) in keys for r in rows for d in offs.tolist())
    assert int((nbr >= 0).sum()) == want


@settings(max_examples=60, deadline=None)
@given(coord_lists, st.sampled_from([2, 4]))
def test_stride_map_properties(rows, ts):
    c = np.asarray(rows, np.int32)
    out, inv = oc.stride_map(c, (ts,) * 3)
    assert np.all(out[:, 1:] % ts == 0)
    assert len({tuple(r) for r in out.tolist()}) == out.shape[0]
    fl = c.copy()
    fl[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
    assert np.array_equal(out[inv], fl)
    first = [np.nonzero(inv == j)[0][0] for j in range(out.shape[0])]
    assert first == sorted(first)                            # first-occurrence order
    # transposed table is the inverse relation of the forward strided table
    fwd = oc.kernel_map_table(c, out, 3, (ts // 2 if ts > 1 else 1,) * 3)
    tr = oc.kernel_map_table(out, c, 3, (ts // 2 if ts > 1 else 1,) * 3, sign=-1)
    for k in range(27):
        o = np.nonzero(fwd[k] >= 0)[0]
        assert np.array_equal(tr[k][fwd[k, o]], o)
    assert int((fwd >= 0).sum()) == int((tr >= 0).sum())


def test_transposed_conv_is_the_adjoint_and_pooling_counts():
    """Self-consistency of the SURVEY 8(f) rank-3 oracle ops: the non-generative transposed convolution with kernel
    W^T is the adjoint of the strided convolution with kernel W on the same pair of maps; average pooling of ones is
    one; the union map keeps the rows of the first operand first."""
    import torch
    from oracle import me_cpu
    rng = np.random.default_rng(0)
    seen, rows = set(), []
    while len(rows) < 500:
        r = (int(rng.integers(2)), *(int(v) for v in rng.integers(-8, 8, 3)))
        if r not in seen:
            seen.add(r)
            rows.append(r)
    c = np.asarray(sorted(rows, key=lambda r: r[0]), dtype=np.int32)
    f = torch.randn(500, 8)
    x = me_cpu.SparseTensor(f, coordinates=torch.from_numpy(c))
    down = me_cpu.MinkowskiConvolution(8, 6, kernel_size=2, stride=2, dimension=3)
    up = me_cpu.MinkowskiConvolutionTranspose(6, 8, kernel_size=2, stride=2, dimension=3)
    with torch.no_grad():
        up.kernel.copy_(down.kernel.transpose(1, 2))
        y = down(x)
        a = torch.randn_like(y.F)
        ua = up(me_cpu.SparseTensor(a, coordinate_map_key=y.coordinate_map_key, coordinate_manager=y.coordinate_manager))
    assert ua.coordinate_map_key == x.coordinate_map_key
    lhs, rhs = float((ua.F * f).sum()), float((a * y.F).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs))
    ones = me_cpu.SparseTensor(torch.ones(500, 2), coordinates=torch.from_numpy(c))
    avg = me_cpu.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)(ones).F
    assert torch.equal(avg, torch.ones_like(avg))
    cm = x.coordinate_manager
    kb = cm.insert(c[::-1][:300] + np.array([0, 1, 0, 0], np.int32), (1, 1, 1), tag="b")
    ukey, ra, rb = cm.union(x.coordinate_map_key, kb)
    assert np.array_equal(cm.coords(ukey)[:500], c) and np.array_equal(ra, np.arange(500))
    assert np.array_equal(cm.coords(ukey)[rb], cm.coords(kb))


def test_input_transform_restatements():
    """``oracle/transforms.py``: the hexagon test agrees with the geometric definition away from the boundary; the
    distance feature is torch's own PairwiseDistance; ``sqrt(fma(dy, dy, dx*dx))`` in fp32 -- the arithmetic the GPU
    kernel uses -- reproduces it to one ulp (bit for bit where torch's CPU kernel fuses the same way)."""
    import torch
    from oracle import transforms as ot
    rng = np.random.default_rng(1)
    xy = rng.uniform(-0.1, 1.1, (20000, 2)).astype(np.float32)
    got = ot.contains_points(ot.HEXAGON, xy)
    # regular hexagon centred at (0.5, 0.5) with circumradius 0.5 and two vertices on the x axis
    dx, dy = np.abs(xy[:, 0].astype(np.float64) - 0.5), np.abs(xy[:, 1].astype(np.float64) - 0.5)
    h = 0.9330127 - 0.5
    geo = (dy < h) & (h * dx + 0.25 * dy < 0.5 * h)
    near = (np.abs(dy - h) < 1e-5) | (np.abs(h * dx + 0.25 * dy - 0.5 * h) < 1e-5)
    assert np.array_equal(got[~near], geo[~near]) and 0.4 < got.mean() < 0.6
    pos = torch.rand(50000, 3)
    f = ot.features(pos)
    d = (pos[:, :2] - torch.tensor([[0.5, 0.5]])) + 1e-6
    a = d.numpy().astype(np.float64)
    fused = np.sqrt((np.float32(a[:, 0] * a[:, 0]).astype(np.float64) + a[:, 1] * a[:, 1]).astype(np.float32)
                    .astype(np.float64)).astype(np.float32)
    assert np.abs(fused - f[:, 2].numpy()).max() <= 1.2e-7 * float(f[:, 2].max())
    assert torch.equal(f[:, 0], torch.ones(50000)) and torch.equal(f[:, 1], pos[:, 2])
    raw = torch.from_numpy(rng.uniform(-15, 15, (1000, 3)).astype(np.float32))
    p, x = ot.test_transform(raw, num=400, perm=torch.randperm)
    assert p.shape[0] <= 400 and float(p[:, 2].min()) >= 0.0 and x.shape == (p.shape[0], 3)


@pytest.mark.parametrize("K", [(7, 7, 7), (3, 3, 3), (8, 3, 2), (1, 5, 5)])
def test_kernel_map_lines_roundtrip(K):
    """The x-line form of a stride-1 map over cell-sorted rows: lines -> table is exact, a hand case pins the packing,
    and unsorted rows are rejected (the form relies on the existing neighbours of a line being consecutive rows)."""
    batch = util.make_points(2, 1500, cfg=11)
    c, _, _, _, _ = util.oracle_quantize(batch, 0.05)
    assert np.all(np.diff(oc.pack_keys(c)) > 0)                                  # (plot, z, y, x), x fastest
    nbr = oc.kernel_map_table(c, c, K, (1, 1, 1))
    lines = oc.kernel_map_lines(nbr, K)
    assert lines.shape == (K[1] * K[2], c.shape[0]) and lines.dtype == np.uint32
    assert np.array_equal(oc.lines_to_table(lines, K), nbr)
    assert int(sum(bin(int(v) & 0xFF).count("1") for v in lines.ravel())) == int((nbr >= 0).sum())


def test_kernel_map_lines_hand_case():
    # one plot, one x-line: cells x = 0, 1, 3 -> rows 0, 1, 2; K = (3, 1, 1)
    c = np.array([[0, 0, 0, 0], [0, 1, 0, 0], [0, 3, 0, 0]], dtype=np.int32)
    nbr = oc.kernel_map_table(c, c, (3, 1, 1), (1, 1, 1))
    lines = oc.kernel_map_lines(nbr, (3, 1, 1))
    # row 0: offsets (-1, 0, +1) -> (none, row 0, row 1): mask 0b110, base 0
    # row 1: (row 0, row 1, none): mask 0b011, base 0;  row 2 (x = 3): (none, row 2, none): mask 0b010, base 2
    assert lines.tolist() == [[(0 << 8) | 0b110, (0 << 8) | 0b011, (2 << 8) | 0b010]]
    shuffled = c[[1, 0, 2]]
    with pytest.raises(AssertionError):
        oc.kernel_map_lines(oc.kernel_map_table(shuffled, shuffled, (3, 1, 1), (1, 1, 1)), (3, 1, 1))


def test_tensor_memory_fragment_channel_order():
    """Layout contract of gather_gemm_ta_kernel (csrc/conv_tc.cu), restated: four lanes q = 0..3 of a row load bytes
    [16 q, 16 q + 16) of the h half and of the l half of a 128-byte operand row chunk ([h(32 bf16) | l(32 bf16)]) and
    store them as their pieces j = 0..3 of tcgen05.st.16x256b.x4, which puts piece j of lane q at tensor-memory columns
    8 j + 2 q + {0, 1} (two bf16 per column).  The k-step s = column // 8 of the MMA then holds, at position
    p = 2 (column % 8) + half, the channel the weight image must carry at the same position: ``ta_channel``."""
    def ta_channel(p):                                   # csrc/conv_tc.cu
        return 8 * ((p & 15) >> 2) + 4 * ((p >> 4) & 1) + (p & 3)

    pos_to_src = {}
    for q in range(4):                                   # lane of the quad
        loads = [(0, 16 * q), (64, 16 * q)]              # (half base, byte offset): v0 from h, v1 from l
        pieces = []                                      # piece j -> (source byte of its first bf16)
        for base, off in loads:
            pieces += [base + off, base + off + 8]       # .xy and .zw of the 16-byte load
        for j, byte in enumerate(pieces):
            for w in range(2):                           # 32-bit word of the piece = one tensor-memory column
                col = 8 * j + 2 * q + w
                for half in range(2):
                    p = 2 * col + half                   # bf16 position in the 64-element K row
                    pos_to_src[p] = (byte + 4 * w + 2 * half) // 2      # source bf16 index: 0..31 = h, 32..63 = l
    assert sorted(pos_to_src) == list(range(64)) and sorted(pos_to_src.values()) == list(range(64))
    for p, src in pos_to_src.items():
        assert (src >= 32) == (p >= 32)                  # k-steps 0, 1 carry h, k-steps 2, 3 carry l
        assert src % 32 == ta_channel(p)                 # same channel order in both halves, as the image builder uses
    # the h k-step s and the l k-step s + 2 hold the same channels: h*H, l*H and h*L pair up step by step
    for p in range(32):
        assert ta_channel(p) == ta_channel(p + 32)
