"""Host-side logic of the data-parallel path on CPU: world_size-2 gloo process group.
Covers the plot sharding and the flat-gradient all-reduce the NCCL path uses on the GPUs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dpcr_agb_b200 import train


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        weights = [13000, 9000, 15000, 7000, 12000, 11000, 14000, 8000]
        mine = train.shard_plots(weights, rank, world)
        # every rank computes a "gradient" from its own plots; the mean over ranks must equal the
        # gradient of the whole batch scaled by 1/world (sum-of-per-plot gradients structure)
        flat = torch.zeros(16)
        for p in mine:
            flat += torch.arange(16, dtype=torch.float32) * weights[p]
        train.allreduce_mean_(flat, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            out.put((gathered, flat.tolist()))
    finally:
        dist.destroy_process_group()


def test_world2_shard_and_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, flat = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    weights = [13000, 9000, 15000, 7000, 12000, 11000, 14000, 8000]
    assert sorted(gathered[0] + gathered[1]) == list(range(8))          # disjoint cover
    assert len(gathered[0]) == len(gathered[1]) == 4                      # equal plot counts (weak scaling)
    loads = [sum(weights[i] for i in g) for g in gathered]
    assert abs(loads[0] - loads[1]) <= 0.05 * sum(weights) / 2            # balanced within 5 %
    expect = [k * sum(weights) / world for k in range(16)]
    assert flat == expect


def test_shard_single_rank_is_identity():
    assert train.shard_plots([3, 1, 2], 0, 1) == [0, 1, 2]
