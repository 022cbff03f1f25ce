"""GPU parity of the input pipeline against the oracle's restatement of the reference's
per-sample transforms (``oracle/transforms.py``): transformed positions, the hexagon mask, MaxPoints selection,
point features, quantised coordinates and the coordinate augmentations -- integers and positions bit-exact."""
import numpy as np
import pytest
import torch

from dpcr_agb_b200.input_pipeline import NFIInputPipeline
from oracle import coords as oc
from oracle import transforms as ot

pytestmark = pytest.mark.gpu


def _raw_plots(num_plots, n_points, seed=0):
    """Raw LiDAR-like points in metres around the plot centre: a disc wider than the hexagon, heights 0..35 m."""
    rng = np.random.default_rng(seed)
    plots = []
    for p in range(num_plots):
        n = n_points + 37 * p
        xy = rng.uniform(-17.0, 17.0, (n, 2))
        z = 50.0 + rng.beta(2, 3, n) * 35.0 + rng.normal(0, 0.02, n)
        plots.append(np.concatenate([xy, z[:, None]], 1).astype(np.float32))
    return plots


def test_pipeline_matches_the_reference_transforms(cuda):
    plots = _raw_plots(3, 5000)
    gen = torch.Generator().manual_seed(3)
    pipe = NFIInputPipeline(max_points=2000)
    ref_pos, ref_x, ranks, orders = [], [], [], []
    for raw in plots:
        pos = ot.start_z_from_zero(ot.move_center(ot.scale_pos(torch.from_numpy(raw), (30.0, 30.0, 40.0))))
        kept, _ = ot.polygon_extend(pos)
        perm = torch.randperm(kept.shape[0], generator=gen)
        rank = torch.empty_like(perm)
        rank[perm] = torch.arange(perm.shape[0])
        ranks.append(rank)
        sel = ot.max_points(kept, 2000, perm)
        ref_pos.append(sel)
        ref_x.append(ot.features(sel))
        orders.append(torch.randperm(sel.shape[0], generator=gen))
    raw_all = torch.from_numpy(np.concatenate(plots)).to(cuda)
    plot_all = torch.cat([torch.full((p.shape[0],), i, dtype=torch.int32) for i, p in enumerate(plots)]).to(cuda)
    # --- steps
    pos, keep = pipe.transform(raw_all, plot_all, 3)
    start = 0
    for raw in plots:
        ref = ot.start_z_from_zero(ot.move_center(ot.scale_pos(torch.from_numpy(raw), (30.0, 30.0, 40.0))))
        assert torch.equal(pos[start:start + raw.shape[0]].cpu(), ref), "transformed positions differ"
        _, mask = ot.polygon_extend(ref)
        assert torch.equal(keep[start:start + raw.shape[0]].cpu().bool(), mask), "hexagon mask differs"
        start += raw.shape[0]
    assert 0.5 < keep.float().mean().item() < 0.95                      # the crop removes a real share of the points
    cpos, cplot, count = pipe.compact(pos, plot_all, keep)
    n_kept = int(count.item())
    assert n_kept == sum(int(r.shape[0]) for r in ranks)
    spos, splot, scount = pipe.max_points_select(cpos, cplot, 3, torch.cat(ranks).to(cuda).int(), count)
    n_sel = int(scount.item())
    assert n_sel == sum(p.shape[0] for p in ref_pos) and all(p.shape[0] == 2000 for p in ref_pos)
    assert torch.equal(spos[:n_sel].cpu(), torch.cat(ref_pos))
    feats = pipe.features(spos, scount)[:n_sel].cpu()
    ref_feats = torch.cat(ref_x)
    assert torch.equal(feats[:, :2], ref_feats[:, :2])
    # torch's CPU PairwiseDistance vectorisation decides where it fuses a multiply-add: allow one ulp
    assert (feats[:, 2] - ref_feats[:, 2]).abs().max().item() <= 1.2e-7 * ref_feats[:, 2].abs().max().item()
    # --- the whole chain incl. the quantiser
    order = torch.cat([o + 2000 * i for i, o in enumerate(orders)]).to(cuda).int()
    vox = pipe(raw_all, plot_all, 3, order=order, max_points_rank=torch.cat(ranks).to(cuda).int())
    c_ref, f_ref, _, _, _ = oc.quantize_batch([p.numpy() for p in ref_pos], [x.numpy() for x in ref_x], 0.0125,
                                              [o.numpy() for o in orders])
    assert np.array_equal(vox["coords"].cpu().numpy(), c_ref)
    got = vox["tensors"][0].cpu().numpy()
    assert np.array_equal(got[:, :2], f_ref[:, :2]) and np.abs(got[:, 2] - f_ref[:, 2]).max() <= 1.2e-7


def test_coordinate_augmentations(cuda):
    plots = _raw_plots(4, 3000, seed=5)
    pipe = NFIInputPipeline()
    raw_all = torch.from_numpy(np.concatenate(plots)).to(cuda)
    plot_all = torch.cat([torch.full((p.shape[0],), i, dtype=torch.int32) for i, p in enumerate(plots)]).to(cuda)
    base = pipe(raw_all, plot_all, 4)
    flips = torch.tensor([[1, 0], [0, 1], [1, 1], [0, 0]])
    shifts = torch.tensor([[3, 99, 0], [0, 0, 0], [57, 1, 20], [7, 7, 7]])
    aug = pipe(raw_all, plot_all, 4, flips=flips, shifts=shifts, resort=False)
    c0, c1 = base["coords"].cpu().numpy(), aug["coords"].cpu().numpy()
    assert aug["index"] is None and base["index"] is not None
    for p in range(4):
        sel = c0[:, 0] == p
        ref = ot.shift_voxels(ot.coords_flip(c0[sel][:, 1:], bool(flips[p, 0]), bool(flips[p, 1])), shifts[p].numpy())
        assert np.array_equal(c1[sel][:, 1:], ref) and np.all(c1[sel][:, 0] == p)
    assert torch.equal(aug["tensors"][0], base["tensors"][0])          # features untouched, same row order

    # resort=True (default): the same voxels in (plot, z, y, x) order with a fresh occupancy index
    srt = pipe(raw_all, plot_all, 4, flips=flips, shifts=shifts)
    c2 = srt["coords"].cpu().numpy()
    order = np.lexsort((c1[:, 1], c1[:, 2], c1[:, 3], c1[:, 0]))
    assert np.array_equal(c2, c1[order])
    assert np.array_equal(srt["row_perm"].cpu().numpy(), order)
    for k in ("pos", "src"):
        assert torch.equal(srt[k].cpu(), aug[k].cpu()[torch.from_numpy(order)])
    assert torch.equal(srt["tensors"][0].cpu(), aug["tensors"][0].cpu()[torch.from_numpy(order)])
    assert srt["index"] is not None
    # static form (no host sync): same rows, given the box of the augmented coordinates
    n_dev = torch.tensor([raw_all.shape[0]], dtype=torch.int32, device=cuda)
    lo, hi = c1[:, 1:].min(0), c1[:, 1:].max(0)
    sta = pipe(raw_all, plot_all, 4, flips=flips, shifts=shifts, bounds=((-80, -80, 0), (80, 80, 100)),
               capacity=c1.shape[0] + 100, n_points_dev=n_dev, aug_bounds=(tuple(lo - 1), tuple(hi + 1)))
    m = int(sta["num_rows"].item())
    assert m == c2.shape[0] and np.array_equal(sta["coords"][:m].cpu().numpy(), c2)
    assert torch.equal(sta["tensors"][0][:m], srt["tensors"][0])


def test_augmented_batch_keeps_the_x_line_stem(cuda):
    """The network on an augmented batch: re-sorted rows + occupancy index (x-line stem, dense maps) give the same
    per-plot predictions as the rows left in the quantiser's order (hash-probed maps, table-driven stem)."""
    from dpcr_agb_b200 import MinkowskiEngine as ME
    from dpcr_agb_b200 import msenet
    from dpcr_agb_b200.MinkowskiEngine import functional as Fn
    plots = _raw_plots(3, 5000, seed=11)
    pipe = NFIInputPipeline()
    raw_all = torch.from_numpy(np.concatenate(plots)).to(cuda)
    plot_all = torch.cat([torch.full((p.shape[0],), i, dtype=torch.int32) for i, p in enumerate(plots)]).to(cuda)
    flips = torch.tensor([[1, 1], [0, 1], [1, 0]])
    shifts = torch.tensor([[5, 0, 2], [0, 9, 0], [1, 1, 1]])
    plain = pipe(raw_all, plot_all, 3, flips=flips, shifts=shifts, resort=False)
    srt = pipe(raw_all, plot_all, 3, flips=flips, shifts=shifts)
    torch.manual_seed(0)
    net = msenet.build(ME, "SENet14", drop_path=0.0).to(cuda).eval()
    x1 = ME.SparseTensor(features=srt["tensors"][0], coordinates=srt["coords"], dense_index=srt["index"])
    km = x1.coordinate_manager.kernel_map(x1.coordinate_map_key, x1.coordinate_map_key, 7)
    assert Fn.lines_path(km, 3, 64)                                   # the stem takes the x-line kernels
    y1 = net(x1)
    y0 = net(ME.SparseTensor(features=plain["tensors"][0], coordinates=plain["coords"]))
    err = ((y1 - y0).abs().max() / y0.abs().max()).item()
    assert err <= 1e-4, err


def test_pipeline_feeds_the_network(cuda):
    """Raw points -> GPU pipeline (static capacities, no host sync) -> MSENet14 forward: finite output per plot."""
    from dpcr_agb_b200 import MinkowskiEngine as ME
    from dpcr_agb_b200 import msenet
    plots = _raw_plots(2, 4000, seed=9)
    pipe = NFIInputPipeline()
    raw_all = torch.from_numpy(np.concatenate(plots)).to(cuda)
    plot_all = torch.cat([torch.full((p.shape[0],), i, dtype=torch.int32) for i, p in enumerate(plots)]).to(cuda)
    dyn = pipe(raw_all, plot_all, 2)
    n_dev = torch.tensor([raw_all.shape[0]], dtype=torch.int32, device=cuda)
    sta = pipe(raw_all, plot_all, 2, bounds=((0, 0, 0), (80, 80, 100)), capacity=dyn["coords"].shape[0] + 256,
               n_points_dev=n_dev)
    m = int(sta["num_rows"].item())
    assert m == dyn["coords"].shape[0] and torch.equal(sta["coords"][:m], dyn["coords"])
    torch.manual_seed(0)
    net = msenet.build(ME, "SENet14", drop_path=0.0).to(cuda).eval()
    y = net(ME.SparseTensor(features=dyn["tensors"][0], coordinates=dyn["coords"], dense_index=dyn["index"]))
    assert y.shape == (2, 2) and torch.isfinite(y).all()
