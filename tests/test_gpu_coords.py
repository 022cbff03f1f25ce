"""GPU parity (bit-exact) of the integer stages against the oracle: quantisation (a1), coordinate maps
(a2, a3), kernel maps (a4).  All calls go through the C ABI (dpcr_agb_b200.lib)."""
import numpy as np
import pytest
import torch

from dpcr_agb_b200 import lib as L
from dpcr_agb_b200.MinkowskiEngine.coordinate_manager import CoordinateManager
from dpcr_agb_b200.quantize import GridSampling3D
from oracle import coords as oc
import b2s_testutil as util

pytestmark = pytest.mark.gpu


def _gpu_quantize(batch, size, dev, with_perm=True, bounds=None):
    gs = GridSampling3D(size, quantize_coords=True, mode="last")
    pos = torch.from_numpy(batch["pos"]).to(dev)
    feats = torch.from_numpy(batch["feats"]).to(dev)
    bidx = torch.from_numpy(batch["batch"]).to(dev)
    order = torch.from_numpy(batch["perm"]).to(dev) if with_perm else None
    return gs(pos, bidx, tensors=(feats,), order=order, bounds=bounds)


@pytest.mark.parametrize("num_plots,n_points,size", [(1, 500, 0.05), (3, 2000, 0.0125), (2, 16000, 0.0125)])
def test_quantize_bit_exact(cuda, num_plots, n_points, size):
    batch = util.make_points(num_plots, n_points)
    c_ref, f_ref, p_ref, s_ref, _ = util.oracle_quantize(batch, size)
    out = _gpu_quantize(batch, size, cuda)
    assert np.array_equal(out["coords"].cpu().numpy(), c_ref)
    assert np.array_equal(out["src"].cpu().numpy().astype(np.int64), s_ref)
    assert np.array_equal(out["tensors"][0].cpu().numpy(), f_ref)
    assert np.array_equal(out["pos"].cpu().numpy(), p_ref)


def test_quantize_identity_order_and_bounds(cuda):
    batch = util.make_points(2, 3000)
    batch_id = dict(batch)
    base = 0
    perm = []
    for b in range(2):
        n = int((batch["batch"] == b).sum())
        perm.append(np.arange(n, dtype=np.int32) + base)
        base += n
    batch_id["perm"] = np.concatenate(perm)
    c_ref, f_ref, _, s_ref, _ = util.oracle_quantize(batch_id, 0.0125)
    out = _gpu_quantize(batch, 0.0125, cuda, with_perm=False, bounds=((0, 0, 0), (80, 80, 110)))
    assert np.array_equal(out["coords"].cpu().numpy(), c_ref)
    assert np.array_equal(out["src"].cpu().numpy().astype(np.int64), s_ref)
    # rows sorted by (plot, z, y, x)  -- size-independent property
    keys = oc.pack_keys(out["coords"].cpu().numpy())
    assert np.all(np.diff(keys) > 0)


def test_quantize_out_of_bounds_raises(cuda):
    batch = util.make_points(1, 400)
    with pytest.raises(L.B2SError):
        _gpu_quantize(batch, 0.0125, cuda, bounds=((0, 0, 0), (10, 10, 10)))


def test_quantize_rounding_half_even(cuda):
    # x.5 cells must round to even exactly like torch.round / np.rint on the fp32 quotient
    size = 0.25
    vals = np.array([0.125, 0.375, 0.625, 0.875, 1.125, -0.125, -0.375, 2.0, 0.3749999], np.float32)
    pos = np.stack([vals, np.zeros_like(vals), np.zeros_like(vals)], 1)
    q_ref = oc.quantize_points(pos, size).astype(np.int32)
    q = torch.empty((len(vals), 3), dtype=torch.int32, device=cuda)
    bnd = torch.empty(6, dtype=torch.int32, device=cuda)
    L.call("b2s_quantize_points", torch.from_numpy(pos).to(cuda), len(vals), None, size, q, bnd)
    assert np.array_equal(q.cpu().numpy(), q_ref)
    assert bnd.tolist() == [int(q_ref[:, 0].min()), 0, 0, int(q_ref[:, 0].max()), 0, 0]


def _manager(coords_np, dev):
    cm = CoordinateManager(D=3, device=dev)
    key, uniq = cm.insert(torch.from_numpy(coords_np).to(dev))
    return cm, key, uniq


def test_insert_unique_and_duplicates(cuda):
    rng = np.random.default_rng(0)
    c = util.random_coords(rng, 5000, nb=3)
    cm, key, uniq = _manager(c, cuda)
    assert uniq is None and cm.num_batches == 3
    assert np.array_equal(cm.coords(key).cpu().numpy(), c)
    # duplicates: first occurrence order
    dup = np.concatenate([c[:100], c[50:300], c[:10]])
    first, inv = oc.unique_first(dup)
    cm2, key2, uniq2 = _manager(dup, cuda)
    assert np.array_equal(uniq2.cpu().numpy(), first)
    assert np.array_equal(cm2.coords(key2).cpu().numpy(), dup[first])


def test_insert_overflow_and_empty(cuda):
    bad = np.array([[0, 40000, 0, 0]], np.int32)
    with pytest.raises(L.B2SError):
        _manager(bad, cuda)
    cm, key, _ = _manager(np.zeros((0, 4), np.int32), cuda)
    assert cm.coords(key).shape[0] == 0


@pytest.mark.parametrize("n,extent", [(3000, 12), (20000, 40)])
def test_stride_maps_bit_exact(cuda, n, extent):
    rng = np.random.default_rng(1)
    c = util.random_coords(rng, n, nb=2, extent=extent)      # includes negative coordinates (floor toward -inf)
    cm, key, _ = _manager(c, cuda)
    cur_ref, cur_key = c, key
    for level in range(4):
        ts = 2 ** (level + 1)
        ref, _ = oc.stride_map(cur_ref, (ts,) * 3)
        cur_key = cm.stride(cur_key, 2)
        got = cm.coords(cur_key).cpu().numpy()
        assert cur_key.tensor_stride == (ts,) * 3
        assert np.array_equal(got, ref), f"level ts={ts}"
        cur_ref = ref


def _check_kmap(cm, in_key, out_key, K, c_in, c_out, ts_in):
    km = cm.kernel_map(in_key, out_key, K)
    step = tuple(ts_in)
    ref = oc.kernel_map_table(c_in, c_out, K, step)
    got = km.nbr.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    # pair-list form (MinkowskiEngine's native layout), canonical order
    i_ref, o_ref, off_ref = oc.table_to_pairs(ref)
    i_g, o_g, off_g = km.pairs()
    assert np.array_equal(off_g.cpu().numpy(), off_ref)
    assert np.array_equal(i_g.cpu().numpy(), i_ref) and np.array_equal(o_g.cpu().numpy(), o_ref)
    return km, ref


def test_kernel_maps_all_hot_path_types(cuda):
    """The 12 distinct maps of an MSENet forward on a real quantised batch."""
    batch = util.make_points(2, 6000)
    c1, _, _, _, _ = util.oracle_quantize(batch, 0.0125)
    cm, k1, _ = _manager(c1, cuda)
    levels = {1: (k1, c1)}
    for ts in (2, 4, 8, 16):
        pk, pc = levels[ts // 2]
        key = cm.stride(pk, 2)
        ref, _ = oc.stride_map(pc, (ts,) * 3)
        assert np.array_equal(cm.coords(key).cpu().numpy(), ref)
        levels[ts] = (key, ref)
    _check_kmap(cm, levels[1][0], levels[1][0], 7, c1, c1, (1, 1, 1))                 # stem k7 s1
    _check_kmap(cm, levels[1][0], levels[2][0], 3, c1, levels[2][1], (1, 1, 1))       # max-pool k3 s2
    for ts in (2, 4, 8, 16):
        key, c = levels[ts]
        km, ref = _check_kmap(cm, key, key, 3, c, c, (ts,) * 3)                        # k3 s1
        assert km.symmetric
        # symmetry property used by dgrad: nbr[k][o] = i  <=>  nbr[K3-1-k][i] = o
        k3 = ref.shape[0]
        for k in (0, 5, 13, 20):
            o = np.nonzero(ref[k] >= 0)[0]
            assert np.array_equal(ref[k3 - 1 - k][ref[k, o]], o)
        if ts < 16:
            nkey, nc = levels[ts * 2]
            km2, ref2 = _check_kmap(cm, key, nkey, 3, c, nc, (ts,) * 3)                # k3 s2
            _check_kmap(cm, key, nkey, 1, c, nc, (ts,) * 3)                            # k1 s2 (downsample)
            inv_ref = oc.kernel_map_table(nc, c, 3, (ts,) * 3, sign=-1)                # transposed map (dgrad)
            assert np.array_equal(km2.inv.cpu().numpy(), inv_ref)
            # transposed table is the exact inverse relation
            for k in range(27):
                o = np.nonzero(ref2[k] >= 0)[0]
                assert np.array_equal(inv_ref[k][ref2[k, o]], o)


def test_kernel_map_even_kernel_and_dilation(cuda):
    rng = np.random.default_rng(3)
    c = util.random_coords(rng, 4000, nb=2, extent=10)
    cm, key, _ = _manager(c, cuda)
    for K, dil in ((2, 1), (3, 2), (5, 1)):
        km = cm.kernel_map(key, key, K, dil)
        ref = oc.kernel_map_table(c, c, K, (dil,) * 3)
        assert np.array_equal(km.nbr.cpu().numpy(), ref)


def test_kernel_map_dense_index_equals_hash(cuda):
    """Kernel maps resolved through the quantiser's occupancy index (bitmap + popcount rank) are bit-identical to
    the hash-probed ones: k7 stride-1 (the stem), k3 stride-2 (max pool), and the transposed table."""
    from dpcr_agb_b200.MinkowskiEngine import coordinate_manager as CM
    batch = util.make_points(3, 6000, cfg=31)
    gs = GridSampling3D(0.02)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) for k, v in batch.items()}
    vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=3)
    tables = {}
    for dense in (True, False):
        CM.USE_DENSE_INDEX = dense
        try:
            cm = CM.CoordinateManager(D=3, device=cuda)
            k1, _ = cm.insert(vox["coords"], dense_index=vox["index"])
            k2 = cm.stride(k1, 2)
            km7 = cm.kernel_map(k1, k1, 7)
            kmp = cm.kernel_map(k1, k2, 3)
            tables[dense] = (km7.nbr.clone(), kmp.nbr.clone(), kmp.inv.clone())
        finally:
            CM.USE_DENSE_INDEX = True
    for a, b in zip(tables[True], tables[False]):
        assert torch.equal(a, b)
    assert (tables[True][0] >= 0).sum() > vox["coords"].shape[0]       # not trivially empty
