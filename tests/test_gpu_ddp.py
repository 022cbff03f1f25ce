"""Data-parallel path on real GPUs: two ranks over NCCL, each with half of the plots, produce after
``Trainer.exchange_gradients`` + the mean folded into the optimiser scale exactly the gradient one rank computes on
the joint batch.  Needs two GPUs (skipped on a one-GPU box; ``tools/gpu_2gpu.sh`` runs it under ``gpurun --gpus 2``).

Batch-norm layers run in eval mode here: the reference keeps BN statistics per replica (``nn.DataParallel``,
trainer.py:149-150), so in training mode the two-rank step is by design NOT the joint-batch step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CFG, N_POINTS, PLOTS = 23, 3000, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _grads(trainer, ME, gs, batch, dev, num_plots):
    """Flat gradient of the mean loss over ``batch`` on this rank (eval-mode network, no exchange)."""
    from dpcr_agb_b200 import train
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in batch.items()}
    vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d["perm"], num_plots=num_plots)
    trainer.opt.zero_grad()
    x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], dense_index=vox["index"])
    pred = trainer.model(x)
    loss = train.reg_loss(pred, d["target"], trainer.center, trainer.scale)
    with trainer.direct_grads():
        loss.backward()
    return float(loss)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from dpcr_agb_b200 import MinkowskiEngine as ME
        from dpcr_agb_b200 import msenet, plots, train
        from dpcr_agb_b200.quantize import GridSampling3D
        torch.manual_seed(0)
        model = msenet.build(ME, "SENet14", drop_path=0.0).to(dev)
        trainer = train.Trainer(model, ME)
        trainer.broadcast_parameters()
        model.eval()
        gs = GridSampling3D(0.02)
        per = PLOTS // world                                           # rank r owns plots [r * per, (r + 1) * per)
        loss_r = _grads(trainer, ME, gs, plots.synth_batch(CFG, rank * per, per, n_points=N_POINTS), dev, per)
        trainer.exchange_gradients()                                   # NCCL sum over the ranks
        mean_grad = (trainer.opt.flat_grad * trainer.opt.grad_scale).clone()   # the 1/world the optimiser kernel applies
        losses = [None] * world
        dist.all_gather_object(losses, loss_r)
        if rank == 0:
            loss_j = _grads(trainer, ME, gs, plots.synth_batch(CFG, 0, PLOTS, n_points=N_POINTS), dev, PLOTS)
            joint = trainer.opt.flat_grad
            err = ((mean_grad - joint).abs().max() / joint.abs().max()).item()
            out.put((err, float(np.mean(losses)), loss_j, float(joint.abs().max())))
    except Exception as e:                                             # fail fast: the parent must not wait for a timeout
        out.put(("error", f"rank {rank}: {type(e).__name__}: {e}", 0.0, 0.0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_gradients_equal_joint_batch_gradients():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    err, loss_mean, loss_joint, gmax = out.get(timeout=240)
    if err == "error":
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
        pytest.fail(str(loss_mean))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert gmax > 0
    assert abs(loss_mean - loss_joint) <= 1e-5 * abs(loss_joint), (loss_mean, loss_joint)
    # same kernels, same operands; only the order of the fp32 sums differs (per-rank partial sums + NCCL ring)
    assert err <= 1e-5, f"two-rank mean gradient differs from the joint-batch gradient by {err:.3e} (relative, max norm)"
