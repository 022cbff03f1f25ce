"""End-to-end GPU parity: MSENet14 / MSENet50 forward + backward through the ME-shaped surface against the
same network over the CPU oracle; fused AdaBelief against the oracle's restatement; full-size
(16 000-point plots, batch of 4) size-independent properties."""
import random

import numpy as np
import pytest
import torch

from dpcr_agb_b200 import MinkowskiEngine as ME
from dpcr_agb_b200 import msenet, plots, train
from dpcr_agb_b200.quantize import GridSampling3D
from oracle import me_cpu
from oracle import train as otrain
import b2s_testutil as util

pytestmark = pytest.mark.gpu

# Per-op parity (tests/test_gpu_conv.py, test_gpu_ops.py) is held to 1e-3 on IDENTICAL inputs, the bar of
# BASELINE.json.  End to end the comparison runs against two precision models of the oracle's convolution
# (oracle/ops.py CONV_PRECISION): "fp32" = the reference's arithmetic, "tf32" = conv operands rounded to TF32
# round-to-nearest with fp32 accumulation = the arithmetic of the tcgen05 kernels.
#   * fp32 SIMT kernels (impl "simt") vs the fp32 oracle, training and eval mode: 1e-3 on outputs, loss, EVERY
#     gradient and the BN running statistics (measured 1e-6 .. 1.3e-4) -- proves the whole pipeline (maps,
#     pooling, SE, batch norm, autograd wiring) exact independently of tensor-core rounding;
#   * tcgen05 kernels, eval mode, vs both models: 1e-3 outputs / 2e-3 gradients (measured 2e-4 .. 6e-4);
#   * tcgen05 kernels, training mode on these SMALL test plots (a few dozen rows in the coarsest maps): batch
#     norm over so few rows amplifies any perturbation ~1e3-fold -- the fp32 oracle and its own tf32 model differ
#     by 3-5 % in some gradients, and the fp32 SIMT path's 1e-7 summation-order noise already shows up as 1e-4.
#     The bounds below are that amplified noise floor, not a kernel tolerance;
#   * tcgen05 kernels, training mode at BASELINE plot size (4 x 16 000 points, 0.0125 grid) -- the configuration
#     the benchmark runs -- where the coarse maps hold hundreds of rows (test_full_size_training_parity).
#   * "tc" = the tcgen05 kernels in their default split-bf16 operand mode (16-17 significant bits per operand):
#     held to the SAME 1e-3 as the SIMT path on these small plots, and to 1e-3 on outputs and EVERY gradient
#     against the fp32 oracle at BASELINE plot size (test_full_size_training_parity_fp32) -- the end-to-end bar of
#     BASELINE.json; "tc_tf32" = the same kernels with TF32 operands (``b2s_set_tuning("precise", 0)``), the
#     faster mode whose end-to-end training error is what single-rounded 10-bit operands give (2e-3 .. 2.5e-2).
E2E_OUT_TOL = {("tc_tf32", False): 1e-3, ("tc_tf32", True): 1e-2, ("simt", False): 1e-3, ("simt", True): 1e-3,
               ("tc", False): 1e-3, ("tc", True): 1e-3}
E2E_GRAD_TOL = {("tc_tf32", False): 2e-3, ("tc_tf32", True): 1e-1, ("simt", False): 1e-3, ("simt", True): 1e-3,
                ("tc", False): 1e-3, ("tc", True): 1e-3}
E2E_BUF_TOL = {"tc_tf32": 1e-2, "simt": 1e-3, "tc": 1e-3}
# absolute slack of the gradient check as a fraction of the largest gradient of the whole model (scalar bias
# gradients are sums with heavy cancellation)
E2E_GRAD_ABS = {("tc_tf32", False): 2e-5, ("tc_tf32", True): 2e-3, ("simt", False): 2e-5, ("simt", True): 2e-5,
                ("tc", False): 2e-5, ("tc", True): 2e-5}


def _report(name, **vals):
    import json
    import os
    path = os.environ.get("B2S_PARITY_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps({"case": name, **vals}) + "\n")


def _pair(name, seed=0, **kw):
    torch.manual_seed(seed)
    ref = msenet.MSENet(me_cpu, name, **kw)
    mine = msenet.MSENet(ME, name, **kw)
    mine.load_state_dict(ref.state_dict())
    return ref, mine


def _grad_errors(mine, ref):
    """Per parameter: (name, abs err, max |oracle grad|); plus the largest oracle gradient overall."""
    rows = []
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        assert n1 == n2
        if p2.grad is None:
            assert p1.grad is None or p1.grad.abs().max().item() == 0.0, n1
            continue
        err = (p1.grad.detach().cpu().double() - p2.grad.double()).abs().max().item()
        rows.append((n1, err, p2.grad.abs().max().item()))
    return rows, max(r[2] for r in rows)


def _grad_check(mine, ref, tol, abs_slack=2e-5):
    rows, gmax = _grad_errors(mine, ref)
    worst, worst_name = 0.0, ""
    for n1, err, pmax in rows:
        if pmax > 1e-3 * gmax and err / pmax > worst:
            worst, worst_name = err / pmax, n1
    if tol is not None:
        for n1, err, pmax in rows:
            bound = tol * pmax + abs_slack * gmax
            assert err <= bound, f"grad of {n1}: abs err {err:.3e} > {bound:.3e} (worst rel {worst:.3e} at {worst_name})"
    return worst, worst_name


CASES = [("simt", "fp32"), ("tc", "bf16x2"), ("tc", "fp32"), ("tc_tf32", "tf32"), ("tc_tf32", "fp32")]


@pytest.mark.parametrize("name,n_points,size", [("SENet14", 2500, 0.04), ("SENet50", 2000, 0.04)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("impl,model", CASES)
def test_msenet_forward_backward(cuda, name, n_points, size, training, impl, model):
    _run_parity(cuda, name, 4, n_points, size, training, impl, model, cfg=7)


def test_full_size_training_parity_fp32(cuda):
    """The end-to-end bar of BASELINE.json where the benchmark runs: MSENet14 TRAINING-mode forward + backward on
    BASELINE-size plots (4 x 16 000 points, grid 0.0125), tcgen05 kernels in their default (split-bf16) operand mode,
    against the FP32 oracle: outputs, loss, BN running statistics and EVERY gradient within 1e-3 -- in the
    max|a-b| / max|b| norm (measured: 1.4e-4 worst) and in SURVEY.md 8(c)'s element-wise norm
    max |a-b| / max(|b|, eps * max|b|) with eps = 0.2 (measured 4e-4 .. 1.1e-3 at eps = 0.1 from run to run -- the
    wgrad partial sums are combined with atomics).  eps cannot usefully go much lower: the fp32 SIMT kernels, whose
    only difference from the oracle is the summation order, already sit at 1e-4 for eps = 0.1 and 9e-4 for eps = 0.01
    on these gradients (tools/parity_diag.py), because training-mode batch norm makes them sums with heavy
    cancellation."""
    _run_parity(cuda, "SENet14", 4, 16000, 0.0125, True, "tc", "fp32", cfg=2, out_tol=1e-3, grad_tol=1e-3,
                buf_tol=1e-3, elementwise_eps=0.2)


def test_full_size_training_parity_tf32_mode(cuda):
    """Same configuration with TF32 operands (``precise`` = 0) against the oracle's tf32 model: the bound single-
    rounded 10-bit operands reach end to end (documented in DESIGN.md; not the default mode)."""
    _run_parity(cuda, "SENet14", 4, 16000, 0.0125, True, "tc_tf32", "tf32", cfg=2, out_tol=2e-3, grad_tol=1e-2,
                buf_tol=1e-3)


def _run_parity(cuda, name, num_plots, n_points, size, training, impl, model, cfg, out_tol=None, grad_tol=None,
                buf_tol=None, elementwise_eps=None):
    from dpcr_agb_b200 import lib
    from dpcr_agb_b200.MinkowskiEngine import functional as Fn
    from oracle import ops as oo
    old_p, oo.CONV_PRECISION = oo.CONV_PRECISION, model
    old_i, Fn.CONV_IMPL = Fn.CONV_IMPL, (1 if impl == "simt" else 0)
    lib.set_tuning("precise", 0 if impl == "tc_tf32" else 1)
    try:
        _parity_body(cuda, name, num_plots, n_points, size, training, impl, model, cfg,
                     out_tol or E2E_OUT_TOL[(impl, training)], grad_tol or E2E_GRAD_TOL[(impl, training)],
                     buf_tol or E2E_BUF_TOL[impl], elementwise_eps)
    finally:
        oo.CONV_PRECISION, Fn.CONV_IMPL = old_p, old_i
        lib.set_tuning("precise", -1)


def _elementwise_check(mine, ref, ym, yr, tol, eps):
    """max over elements of |a-b| / max(|b|, eps * max|b|), per tensor; parameters whose oracle
    gradient is numerically zero (a bias in front of a training-mode batch norm) are measured against 1e-3 of the
    largest gradient component of the model instead of their own, meaningless, scale."""
    def err(a, b, floor=0.0):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        den = torch.clamp(b.abs(), min=max(eps * b.abs().max().item(), floor, 1e-300))
        return ((a - b).abs() / den).max().item()
    worst, worst_name = err(ym, yr), "output"
    assert worst <= tol, f"output: element-wise error {worst:.3e} > {tol:.1e}"
    gmax = max(p.grad.abs().max().item() for p in ref.parameters() if p.grad is not None)
    for (n1, p1), (_, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        if p2.grad is None:
            continue
        e = err(p1.grad, p2.grad, floor=1e-3 * gmax)
        assert e <= tol, f"grad of {n1}: element-wise error {e:.3e} > {tol:.1e}"
        if e > worst:
            worst, worst_name = e, n1
    return worst, worst_name


def _parity_body(cuda, name, num_plots, n_points, size, training, impl, model, cfg, out_tol, grad_tol, buf_tol,
                 elementwise_eps=None):
    batch = util.make_points(num_plots, n_points, cfg=cfg)
    c, f, _, _, _ = util.oracle_quantize(batch, size)
    ref, mine = _pair(name, drop_path=0.2)
    mine = mine.to(cuda)
    ref.train(training)
    mine.train(training)
    target = torch.from_numpy(batch["target"])
    center, scale = torch.tensor([107.0, 200.0]), torch.tensor([103.0, 194.0])

    random.seed(11)
    yr = ref(me_cpu.SparseTensor(torch.from_numpy(f), coordinates=torch.from_numpy(c)))
    lr = otrain.reg_loss(yr, target, center, scale)
    lr.backward()
    random.seed(11)
    ym = mine(ME.SparseTensor(features=torch.from_numpy(f), coordinates=torch.from_numpy(c), device=cuda))
    lm = train.reg_loss(ym, target.to(cuda), center.to(cuda), scale.to(cuda))
    lm.backward()
    tag = f"{name}-n{n_points}-train{training}-{impl}-vs-{model}"
    worst, worst_name = _grad_check(mine, ref, None)
    _report(tag, rel_err_out=util.rel_err(ym, yr), rel_err_loss=util.rel_err(lm, lr), worst_rel_err_grad=worst,
            worst_grad=worst_name)
    util.assert_close(ym, yr, tol=out_tol, what=f"{name} output")
    util.assert_close(lm, lr, tol=out_tol, what=f"{name} loss")
    _grad_check(mine, ref, grad_tol, E2E_GRAD_ABS[(impl, training)] if n_points < 16000 else
                (1e-5 if impl == "tc" else 1e-4))
    if elementwise_eps is not None:
        for eps in (0.05, 0.1):                       # reported, not asserted
            ew, ew_name = _elementwise_check(mine, ref, ym, yr, 1e9, eps)
            _report(tag + "-elementwise", eps=eps, worst=ew, worst_tensor=ew_name)
        ew, ew_name = _elementwise_check(mine, ref, ym, yr, grad_tol, elementwise_eps)
        _report(tag + "-elementwise", eps=elementwise_eps, worst=ew, worst_tensor=ew_name)
    if training:   # BN running statistics followed the same batches
        for (n1, b1), (n2, b2) in zip(mine.named_buffers(), ref.named_buffers()):
            if b2.dtype.is_floating_point:
                util.assert_close(b1, b2, tol=buf_tol, what=f"buffer {n1}")
            else:
                assert int(b1) == int(b2)


def test_state_dict_round_trip_and_names(cuda):
    ref, mine = _pair("SENet14")
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    assert "blocks.0.0.conv.kernel" in mine.state_dict() and "blocks.2.0.downsample.1.bn.weight" in mine.state_dict()
    assert tuple(mine.state_dict()["blocks.0.0.conv.kernel"].shape) == (343, 3, 64)
    assert tuple(mine.state_dict()["blocks.2.0.downsample.0.kernel"].shape) == (1, 64, 128)


def test_fused_adabelief_matches_oracle(cuda):
    torch.manual_seed(0)
    shapes = [(27, 8, 16), (1, 16), (16,), (5, 7)]
    ps_ref = [torch.randn(s) for s in shapes]
    ps_gpu = [p.clone().to(cuda).requires_grad_() for p in ps_ref]
    for p in ps_ref:
        p.requires_grad_()
    oref = otrain.AdaBelief(ps_ref, lr=5e-3, weight_decay=1e-2)
    ogpu = train.FlatAdaBelief(ps_gpu, lr=5e-3, weight_decay=1e-2, grad_clip=100.0)
    for step in range(12):        # crosses the num_sma >= 5 switch (step 6)
        for pr, pg in zip(ps_ref, ps_gpu):
            g = torch.randn(pr.shape) * (200.0 if step == 3 else 1.0)     # step 3 exercises the clip
            pr.grad = g.clamp(-100.0, 100.0)
            pg.grad.copy_(g.to(cuda))
        oref.step()
        ogpu.step()
        for pr, pg in zip(ps_ref, ps_gpu):
            util.assert_close(pg, pr, tol=2e-6, what=f"param after step {step}")
    # inf skip (GradScaler semantics)
    before = ogpu.flat_param.clone()
    ogpu.flat_grad[3] = float("inf")
    ogpu.step(check_inf=True)
    assert torch.equal(before, ogpu.flat_param)


def test_fused_adabelief_parameter_groups(cuda):
    """Head / backbone parameter groups with their own lr and weight decay (the reference's
    ``MinkowskiBaselineModel.get_parameter_list``, models/instance/minkowski.py:54-65): every group is a segment of the
    flat buffer updated with its own hyper-parameters; checked against one oracle optimiser per group, eager
    (host hyper-parameters) and through the device-side hyper block the captured step reads, with a scheduled lr."""
    torch.manual_seed(1)
    shapes = [(27, 8, 16), (16,), (5, 7), (7,)]
    ps_ref = [torch.randn(s) for s in shapes]
    ps_gpu = [p.clone().to(cuda).requires_grad_() for p in ps_ref]
    for p in ps_ref:
        p.requires_grad_()
    head_ref, back_ref = ps_ref[2:], ps_ref[:2]
    o_head = otrain.AdaBelief(head_ref, lr=1e-2, weight_decay=0.0)
    o_back = otrain.AdaBelief(back_ref, lr=5e-3, weight_decay=1e-2)
    ogpu = train.FlatAdaBelief([{"params": ps_gpu[2:], "lr": 1e-2, "weight_decay": 0.0}, {"params": ps_gpu[:2]}],
                               lr=5e-3, weight_decay=1e-2, grad_clip=100.0)
    assert [g["numel"] for g in ogpu.groups] == [42, 27 * 8 * 16 + 16]
    order = ps_ref[2:] + ps_ref[:2]                      # flat-buffer order: group by group
    for step in range(8):
        factor = 1.0 - 0.1 * step                        # a scheduler scales every group's lr by the same factor
        o_head.lr, o_back.lr, ogpu.lr = 1e-2 * factor, 5e-3 * factor, 5e-3 * factor
        for pr, pg in zip(order, ogpu.params):
            g = torch.randn(pr.shape)
            pr.grad = g.clone()
            pg.grad.copy_(g.to(cuda))
        o_head.step()
        o_back.step()
        if step % 2 == 0:
            ogpu.step()
        else:                                            # the captured step's form: hyper-parameters from device memory
            ogpu.upload_hyper()
            ogpu.step_from_device(skip_flag=False)
        for pr, pg in zip(order, ogpu.params):
            util.assert_close(pg, pr, tol=2e-6, what=f"param after step {step}")
    sd = ogpu.state_dict()
    assert [len(g["params"]) for g in sd["param_groups"]] == [2, 2] and sd["param_groups"][0]["weight_decay"] == 0.0


def test_trainer_steps_reduce_loss(cuda):
    """Three optimisation steps on one small batch through the public Trainer; loss must go down and stay
    finite (the functional smoke of the whole path: quantise -> hash -> maps -> convs -> optimiser)."""
    batch = util.make_points(4, 2000, cfg=9)
    gs = GridSampling3D(0.0125)
    out = gs(torch.from_numpy(batch["pos"]).to(cuda), torch.from_numpy(batch["batch"]).to(cuda),
             tensors=(torch.from_numpy(batch["feats"]).to(cuda),), order=torch.from_numpy(batch["perm"]).to(cuda))
    torch.manual_seed(0)
    model = msenet.build(ME, "SENet14", drop_path=0.0).to(cuda)
    tr = train.Trainer(model, ME, lr=2e-4)
    before = tr.opt.flat_param.clone()
    target = torch.from_numpy(batch["target"]).to(cuda)
    losses = [float(tr.step(out["coords"], out["tensors"][0], target)) for _ in range(6)]
    assert all(np.isfinite(losses)), losses
    assert min(losses[1:]) < losses[0], losses
    assert not torch.equal(before, tr.opt.flat_param)


def test_full_size_properties(cuda):
    """BASELINE-size plots (16 000 points, 0.0125 grid): properties that need no oracle run."""
    batch = util.make_points(4, 16000, cfg=2)
    gs = GridSampling3D(0.0125)
    out = gs(torch.from_numpy(batch["pos"]).to(cuda), torch.from_numpy(batch["batch"]).to(cuda),
             tensors=(torch.from_numpy(batch["feats"]).to(cuda),), order=torch.from_numpy(batch["perm"]).to(cuda))
    coords = out["coords"]
    # (1) quantise is idempotent on its own representatives and sorted by (plot, z, y, x)
    c = coords.cpu().numpy().astype(np.int64)
    key = ((c[:, 0] * 4096 + c[:, 3]) * 4096 + c[:, 2]) * 4096 + c[:, 1]
    assert np.all(np.diff(key) > 0)
    # (2) every voxel's representative maps back to the voxel
    q = np.rint(batch["pos"][out["src"].cpu().numpy()] / np.float32(0.0125)).astype(np.int32)
    assert np.array_equal(q, coords.cpu().numpy()[:, 1:])
    x = ME.SparseTensor(features=out["tensors"][0], coordinates=coords)
    cm = x.coordinate_manager
    k1 = x.coordinate_map_key
    k2 = cm.stride(k1, 2)
    # (3) strided map: floor of every fine row exists exactly once in the coarse map
    fine = cm.coords(k1)
    coarse = cm.coords(k2)
    fl = fine.clone()
    fl[:, 1:] = torch.div(fine[:, 1:], 2, rounding_mode="floor") * 2
    assert torch.unique(fl, dim=0).shape[0] == coarse.shape[0]
    assert torch.unique(coarse, dim=0).shape[0] == coarse.shape[0]
    # (4) kernel-map symmetry and centre tap on the stem map (343 offsets)
    km = cm.kernel_map(k1, k1, 7)
    n = km.n_out
    assert torch.equal(km.nbr[171], torch.arange(n, device=cuda, dtype=torch.int32))      # centre offset
    for k in (0, 100, 250):
        o = torch.nonzero(km.nbr[k] >= 0).squeeze(1)
        assert torch.equal(km.nbr[342 - k][km.nbr[k][o].long()].long(), o)
    # (5) conv linearity on the full-size map: conv(a x1 + x2) = a conv(x1) + conv(x2)
    conv = ME.MinkowskiConvolution(3, 64, kernel_size=7, stride=1, dimension=3, bias=False).to(cuda)
    f1, f2 = out["tensors"][0], torch.randn_like(out["tensors"][0])
    y1 = conv(ME.SparseTensor(f1, coordinate_map_key=k1, coordinate_manager=cm)).F
    y2 = conv(ME.SparseTensor(f2, coordinate_map_key=k1, coordinate_manager=cm)).F
    y3 = conv(ME.SparseTensor(2.0 * f1 + f2, coordinate_map_key=k1, coordinate_manager=cm)).F
    util.assert_close(y3, 2.0 * y1 + y2, tol=2e-3, what="linearity")


@pytest.mark.parametrize("name", ["SENet14", "SENet50"])
def test_tf32_twins_match_the_rounding_pass_and_remove_it(cuda, name):
    """Producers that write the TF32 operand of the next convolution themselves (C ABI ``*_tf32`` outputs) must give
    the results of the separate b2s_round_tf32 pass -- same rounding, same operands; what is left is the summation
    order of the split-K / wgrad atomics, 1e-7-level noise, hence eval-mode batch norm and a 1e-5 / 1e-4 bound --
    while the number of rounding launches drops to the few convolutions whose input has no fused producer."""
    from dpcr_agb_b200 import lib
    from dpcr_agb_b200.MinkowskiEngine import functional as Fn
    batch = util.make_points(2, 2500, cfg=2)
    c, f, _, _, _ = util.oracle_quantize(batch, 0.04)
    torch.manual_seed(3)
    model = msenet.MSENet(ME, name, drop_path=0.0).to(cuda).eval()
    target = torch.from_numpy(batch["target"]).to(cuda)
    center, scale = torch.tensor([107.0, 200.0], device=cuda), torch.tensor([103.0, 194.0], device=cuda)
    runs = {}
    old = Fn.TWINS
    try:
        for twins in (False, True):
            Fn.TWINS = twins
            model.zero_grad(set_to_none=True)
            for b in model.buffers():          # same BN running statistics at the start of both runs
                if b.dtype.is_floating_point:
                    b.fill_(0.5)
            lib.profile_start(["b2s_round_tf32"])
            y = model(ME.SparseTensor(features=torch.from_numpy(f), coordinates=torch.from_numpy(c), device=cuda))
            loss = train.reg_loss(y, target, center, scale)
            loss.backward()
            rounds = sum(n for n, _ in lib.profile_stop().values())
            runs[twins] = (y.detach().clone(), [p.grad.detach().clone() for p in model.parameters()], rounds)
    finally:
        Fn.TWINS = old
    util.assert_close(runs[True][0], runs[False][0], tol=1e-5, what="output with / without twins")
    gmax = max(g.abs().max().item() for g in runs[False][1])
    for g1, g0 in zip(runs[True][1], runs[False][1]):
        # summation-order noise scales with the summands, not with the (possibly cancelling) sum
        assert (g1 - g0).abs().max().item() <= 1e-3 * g0.abs().max().item() + 1e-5 * gmax
    assert runs[True][2] < runs[False][2] // 2, f"rounding launches {runs[False][2]} -> {runs[True][2]}"
    _report(f"{name}-tf32-twins", rounding_launches_without=runs[False][2], rounding_launches_with=runs[True][2])


@pytest.mark.parametrize("name", ["SENet14", "SENet50"])
def test_fused_se_tail_matches_the_unfused_chain(cuda, name):
    """``ME.fused_se_tail`` (per-plot mean, excitation MLP, gate x drop-path scale, residual add, GELU and their
    backward as six kernels) against the same block tail composed of the individual ME ops: outputs and every
    gradient, batch norm in eval mode (so that only summation order differs), drop path active."""
    batch = util.make_points(3, 2500, cfg=2)
    c, f, _, _, _ = util.oracle_quantize(batch, 0.04)
    torch.manual_seed(5)
    model = msenet.MSENet(ME, name, drop_path=0.3).to(cuda).eval()
    blocks = [b for st in model.blocks[1:] for b in st]
    assert all(b._se_tail is not None for b in blocks)
    for b in blocks:
        b.drop_path.train()
    target = torch.from_numpy(batch["target"]).to(cuda)
    center, scale = torch.tensor([107.0, 200.0], device=cuda), torch.tensor([103.0, 194.0], device=cuda)
    tails = [b._se_tail for b in blocks]
    runs = {}
    try:
        for fused in (False, True):
            for b, t in zip(blocks, tails):
                b._se_tail = t if fused else None
            model.zero_grad(set_to_none=True)
            random.seed(23)
            y = model(ME.SparseTensor(features=torch.from_numpy(f), coordinates=torch.from_numpy(c), device=cuda))
            train.reg_loss(y, target, center, scale).backward()
            runs[fused] = (y.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters()})
    finally:
        for b, t in zip(blocks, tails):
            b._se_tail = t
    # u * (gate * keep) instead of (u * gate) * keep: one-ulp differences, which the TF32 rounding of the next
    # convolution's operands turns into occasional 5e-4 flips of single operands -> 1e-5 .. 1e-4 downstream
    util.assert_close(runs[True][0], runs[False][0], tol=2e-4, what="output, fused vs unfused SE tail")
    gmax = max(g.abs().max().item() for g in runs[False][1].values())
    for n, g0 in runs[False][1].items():
        err = (runs[True][1][n] - g0).abs().max().item()
        assert err <= 2e-3 * g0.abs().max().item() + 1e-4 * gmax, f"{n}: {err:.3e}"


def test_direct_param_grads_refuses_a_second_backward(cuda):
    """In direct mode the backward kernels OVERWRITE ``param.grad`` (views of the trainer's zeroed flat buffer): a second
    backward before ``zero_grad`` would silently discard the first gradient, so it raises; after ``zero_grad`` the
    next step runs normally."""
    from dpcr_agb_b200 import train
    torch.manual_seed(0)
    model = msenet.build(ME, "SENet14", drop_path=0.0).to(cuda)
    tr = train.Trainer(model, ME, lr=1e-4)
    b = util.make_points(2, 1500, cfg=13)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda) for k, v in b.items()}
    vox = GridSampling3D(0.04)(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=2)

    def loss():
        x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"])
        return train.reg_loss(model(x), d["target"], tr.center, tr.scale)

    model.train()
    tr.opt.zero_grad()
    with tr.direct_grads():
        loss().backward()
        with pytest.raises(RuntimeError, match="already written in place"):
            loss().backward()
    tr.opt.zero_grad()
    with tr.direct_grads():
        loss().backward()
    assert torch.isfinite(tr.opt.flat_grad).all() and tr.opt.flat_grad.abs().max() > 0
