"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/b200sparse.h
declares; the Python module exports every ME.* name the reference touches; the UNCHANGED
reference networks import and build over it with state-dict keys identical to our restatement."""
import os
import re
import sys
import types
import warnings

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/torch-points3d"


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    return sorted(set(re.findall(r"B2S_API\s+[\w\s\*]+?\b(b2s_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from dpcr_agb_b200 import lib
    names = _declared_symbols()
    assert len(names) >= 30
    handle = lib.load()                       # raises if the .so is missing: there is no fallback
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert handle.b2s_version() >= 100


def test_product_path_has_no_cpu_fallback():
    from dpcr_agb_b200 import MinkowskiEngine as ME
    from dpcr_agb_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("CPU-only assertion")
    with pytest.raises(RuntimeError):
        ME.SparseTensor(features=torch.zeros(2, 3), coordinates=torch.zeros(2, 4, dtype=torch.int32))
    with pytest.raises(lib.B2SError):
        lib.call("b2s_gelu_fwd", torch.zeros(4), 1, None, 4, torch.zeros(4))
    # nothing of the product imports (or loads by path) the oracle: every .py file of the package is scanned
    import glob
    import re
    pkg = os.path.dirname(os.path.abspath(lib.__file__))
    files = glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True)
    assert len(files) >= 15
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b|from\s+\.+\s*oracle\b)|import_module\([\"']oracle|__import__\([\"']oracle",
                     re.M)
    for f in files:
        assert not pat.search(open(f).read()), f"{f} imports the oracle"


ME_NAMES = """SparseTensor MinkowskiConvolution MinkowskiConvolutionTranspose MinkowskiMaxPooling MinkowskiAvgPooling
MinkowskiSumPooling MinkowskiAvgUnpooling MinkowskiGlobalPooling MinkowskiGlobalSumPooling MinkowskiGlobalAvgPooling
MinkowskiGlobalMaxPooling MinkowskiBroadcastMultiplication MinkowskiLinear MinkowskiBatchNorm MinkowskiInstanceNorm
MinkowskiDropout MinkowskiReLU MinkowskiSigmoid MinkowskiNetwork RegionType KernelGenerator cat utils
MinkowskiNormalization MinkowskiNonlinearity CoordinateManager""".split()


def test_me_surface_names():
    from dpcr_agb_b200 import MinkowskiEngine as ME
    for n in ME_NAMES:
        assert hasattr(ME, n), n
    for n in "MinkowskiGELU MinkowskiReLU MinkowskiCELU MinkowskiSiLU MinkowskiELU MinkowskiSigmoid MinkowskiTanh MinkowskiSinusoidal".split():
        assert hasattr(ME.MinkowskiNonlinearity, n), n
    for n in "MinkowskiBatchNorm MinkowskiInstanceNorm".split():
        assert hasattr(ME.MinkowskiNormalization, n), n
    assert {m.name for m in ME.RegionType} >= {"HYPER_CUBE", "HYPER_CROSS", "CUSTOM"}
    assert callable(ME.utils.kaiming_normal_)
    conv = ME.MinkowskiConvolution(3, 64, kernel_size=7, stride=1, dimension=3, bias=True)
    assert tuple(conv.kernel.shape) == (343, 3, 64) and tuple(conv.bias.shape) == (1, 64)
    assert tuple(ME.MinkowskiConvolution(64, 128, kernel_size=1, stride=2, dimension=3).kernel.shape) == (1, 64, 128)
    assert tuple(ME.MinkowskiConvolution(64, 256, kernel_size=1, dimension=3).kernel.shape) == (64, 256)  # use_mm
    assert isinstance(ME.MinkowskiBatchNorm(8).bn, torch.nn.BatchNorm1d)
    assert isinstance(ME.MinkowskiLinear(8, 2).linear, torch.nn.Linear)
    # SparseTensor must pass through torch.cuda.amp.custom_fwd untouched (senet_block.py:46)
    import collections.abc
    assert not issubclass(ME.SparseTensor, collections.abc.Mapping) and not hasattr(ME.SparseTensor, "__iter__")


def _stub_reference_imports():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
    stub("omegaconf", OmegaConf=type("OmegaConf", (), {}), DictConfig=dict, ListConfig=list)
    stub("omegaconf.listconfig", ListConfig=list)
    stub("omegaconf.dictconfig", DictConfig=dict)
    stub("matplotlib")
    stub("matplotlib.pyplot")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
@pytest.mark.parametrize("backend", ["product", "oracle"])
def test_unchanged_reference_networks_build_over_our_module(backend):
    """`import MinkowskiEngine as ME` inside the reference resolves to our module; SENet14 / SENet50 construct,
    init_weights runs (touches .kernel/.bias/.bn/.linear), and the state-dict keys equal our restatement's."""
    import dpcr_agb_b200
    from dpcr_agb_b200 import msenet
    from oracle import me_cpu
    for k in [k for k in sys.modules if k.startswith("torch_points3d") or k.startswith("MinkowskiEngine")]:
        del sys.modules[k]
    ME = dpcr_agb_b200.install() if backend == "product" else me_cpu.install()
    _stub_reference_imports()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from torch_points3d.modules.MinkowskiEngine import SENet, initialize_minkowski_unet
    try:
        for name, nparam in (("SENet14", 14447870), ("SENet50", 48760626)):
            ref = initialize_minkowski_unet(name, 3, 2, activation="gelu", first_stride=1, global_pool="sum",
                                            bias=True, bn_momentum=0.1, norm_type="bn", dropout=0.0, drop_path=0.01)
            assert isinstance(ref, getattr(SENet, name))
            mine = msenet.MSENet(ME, name, drop_path=0.01, separate_head=False)
            a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
            b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
            assert a == b
            assert sum(p.numel() for p in ref.parameters()) == nparam
            mine.load_state_dict(ref.state_dict())
    finally:
        for k in [k for k in sys.modules if k.startswith("torch_points3d") or k.startswith("MinkowskiEngine")]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_reference_senet14_equals_restatement_on_oracle():
    """Forward of the UNCHANGED reference SENet14 == forward of dpcr_agb_b200.msenet on the same (oracle)
    backend, same weights, same DropPath draws: pins our network restatement to the reference definition."""
    import random
    import numpy as np
    from dpcr_agb_b200 import msenet, plots
    from oracle import coords as oc
    from oracle import me_cpu
    for k in [k for k in sys.modules if k.startswith("torch_points3d") or k.startswith("MinkowskiEngine")]:
        del sys.modules[k]
    me_cpu.install()
    _stub_reference_imports()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from torch_points3d.modules.MinkowskiEngine import SENet
        try:
            torch.manual_seed(0)
            ref = SENet.SENet14(in_channels=3, out_channels=2, activation="gelu", first_stride=1, global_pool="sum",
                                drop_path=0.3, D=3)
            mine = msenet.MSENet(me_cpu, "SENet14", drop_path=0.3, separate_head=False)
            mine.load_state_dict(ref.state_dict())
            b = plots.synth_batch(0, 0, 2, n_points=800)
            c, f, _, _, _ = oc.quantize_batch([b["pos"][b["batch"] == i] for i in range(2)],
                                              [b["feats"][b["batch"] == i] for i in range(2)], 0.05)
            for training in (False, True):
                ref.train(training)
                mine.train(training)
                random.seed(4)
                yr = ref(me_cpu.SparseTensor(torch.from_numpy(f), coordinates=torch.from_numpy(c))).F
                random.seed(4)
                ym = mine(me_cpu.SparseTensor(torch.from_numpy(f), coordinates=torch.from_numpy(c))).F
                assert torch.equal(yr, ym)
        finally:
            for k in [k for k in sys.modules if k.startswith("torch_points3d") or k.startswith("MinkowskiEngine")]:
                del sys.modules[k]


def test_block_output_forms_follow_the_consumers():
    """Host logic of the TF32-twin scheme (dpcr_agb_b200/msenet.py): a block whose output is consumed by convolutions
    only (the next block has a downsample convolution) writes it TF32-rounded (2), one whose output also feeds an
    identity residual writes the plain result plus the twin (1), the last block writes the plain result (0)."""
    from dpcr_agb_b200 import MinkowskiEngine as ME
    from dpcr_agb_b200 import msenet
    flags = {}
    for name in ("SENet14", "SENet50"):
        m = msenet.MSENet(ME, name)
        blocks = [b for st in m.blocks[1:] for b in st]
        flags[name] = [b.out_tf32 for b in blocks]
        for blk, nxt in zip(blocks, blocks[1:]):
            has_down = not isinstance(nxt.downsample, torch.nn.Identity)
            assert blk.out_tf32 == (2 if has_down else 1)
        assert all(b._se_tail is not None for b in blocks)          # channel counts of both nets suit the fused tail
    assert flags["SENet14"] == [2, 2, 2, 0]
    assert flags["SENet50"] == [1, 1, 2, 1, 1, 1, 2, 1, 1, 1, 1, 1, 2, 1, 1, 0]


def test_step_contexts_restore_their_flags():
    """``direct_param_grads`` / ``deferred_bn_counters`` are process-wide switches used around one step: they must
    nest and restore, and the deferred counters must be applied exactly once on exit."""
    from dpcr_agb_b200.MinkowskiEngine import functional as Fn
    from dpcr_agb_b200.MinkowskiEngine import modules as M
    assert Fn.DIRECT_PARAM_GRADS is False
    with Fn.direct_param_grads():
        assert Fn.DIRECT_PARAM_GRADS is True
        with Fn.direct_param_grads():
            assert Fn.DIRECT_PARAM_GRADS is True
        assert Fn.DIRECT_PARAM_GRADS is True
    assert Fn.DIRECT_PARAM_GRADS is False
    counters = [torch.zeros((), dtype=torch.long) for _ in range(3)]
    assert M.DEFERRED_BN_COUNTERS is None
    with M.deferred_bn_counters():
        for t in counters:
            M.DEFERRED_BN_COUNTERS.append(t)
        assert all(int(t) == 0 for t in counters)
    assert M.DEFERRED_BN_COUNTERS is None and all(int(t) == 1 for t in counters)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_reference_pointnet_resnet_minkunet_run_over_the_oracle():
    """SURVEY.md 8(f) ranks 3-4 on the oracle side: the UNCHANGED reference ``MinkowskiPointNet`` equals our
    restatement (state-dict keys and forward), and the unchanged ``ResNet14`` (k5s2 conv, k2s2 average pooling, k3s3
    conv, global max pooling) and ``MinkUNet14A`` (k2s2 transposed convolutions onto the encoder maps, ``ME.cat``)
    of ``networks.py`` / ``api_modules`` run forward + backward over it."""
    import numpy as np
    from dpcr_agb_b200 import plots, pointnet
    from oracle import coords as oc
    from oracle import me_cpu
    for k in [k for k in sys.modules if k.startswith("torch_points3d") or k.startswith("MinkowskiEngine")]:
        del sys.modules[k]
    me_cpu.install()
    _stub_reference_imports()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from torch_points3d.modules.MinkowskiEngine import PointNet, networks
        try:
            b = plots.synth_batch(0, 0, 2, n_points=600)
            c, f, _, _, _ = oc.quantize_batch([b["pos"][b["batch"] == i] for i in range(2)],
                                              [b["feats"][b["batch"] == i] for i in range(2)], 0.05)
            x6 = np.concatenate([c[:, 1:].astype(np.float32) * 0.05, f], 1)
            torch.manual_seed(0)
            ref = PointNet.MinkowskiPointNet(3, 2, activation="gelu", global_pool="max", embedding_channel=256)
            mine = pointnet.MinkowskiPointNet(me_cpu, 3, 2, activation="gelu", global_pool="max", embedding_channel=256)
            assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
                {k: tuple(v.shape) for k, v in mine.state_dict().items()}
            mine.load_state_dict(ref.state_dict())
            yr = ref(me_cpu.SparseTensor(torch.from_numpy(x6), coordinates=torch.from_numpy(c))).F
            ym = mine(me_cpu.SparseTensor(torch.from_numpy(x6), coordinates=torch.from_numpy(c))).F
            assert torch.equal(yr, ym)
            for cls in (networks.ResNet14, networks.MinkUNet14A):
                torch.manual_seed(1)
                net = cls(3, 4, D=3)
                x = me_cpu.SparseTensor(torch.from_numpy(f).clone().requires_grad_(), coordinates=torch.from_numpy(c))
                y = net(x)
                assert y.F.shape[1] == 4 and torch.isfinite(y.F).all()
                if cls is networks.MinkUNet14A:
                    assert y.F.shape[0] == c.shape[0] and y.coordinate_map_key == x.coordinate_map_key
                y.F.square().sum().backward()
                assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
        finally:
            for k in [k for k in sys.modules if k.startswith("torch_points3d") or k.startswith("MinkowskiEngine")]:
                del sys.modules[k]


def test_checkpoint_layout_round_trip(tmp_path):
    """``dpcr_agb_b200.checkpoint`` writes the dict layout of ``metrics/model_checkpoint.py:24-61`` (models / optimizer /
    schedulers / stats / run_config) with the backbone under ``model.`` as ``MinkowskiBaselineModel`` saves it, and
    loads both layouts back."""
    from dpcr_agb_b200 import MinkowskiEngine as ME
    from dpcr_agb_b200 import checkpoint, msenet
    torch.manual_seed(0)
    a = msenet.MSENet(ME, "SENet14")
    path = str(tmp_path / "SENet14.pt")
    obj = checkpoint.save(path, a, run_config={"model_name": "SENet14"}, extra_models={"best_loss": a.state_dict()})
    assert set(obj) >= {"run_config", "models", "stats", "optimizer", "schedulers", "grad_scale", "dataset_properties"}
    assert set(obj["models"]) == {"latest", "best_loss"}
    assert all(k.startswith("model.") for k in obj["models"]["latest"])
    assert "model.blocks.0.0.conv.kernel" in obj["models"]["latest"] and "model.final.linears.1.bias" in obj["models"]["latest"]
    torch.manual_seed(1)
    b = msenet.MSENet(ME, "SENet14")
    checkpoint.load(path, b, weight_name="best_miou")            # unknown name -> latest (model_checkpoint.py:237-243)
    for (k, va), (_, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(va, vb), k
    torch.save({"models": {"latest": a.state_dict()}}, path)      # a file without the model. prefix loads too
    checkpoint.load(path, b)
