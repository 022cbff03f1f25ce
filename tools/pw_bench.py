"""Micro-benchmark of the column reductions (batch-norm statistics, batch-norm backward sums, bias-gradient column
sum) at the MSENet14 layer shapes of one 32-plot batch, replayed from a CUDA graph (no launch gaps) over rotating
buffers (cold for the large layers), for every knob combination in CFGS.  GPU box only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dpcr_agb_b200 import lib as L

dev = torch.device("cuda:0")
L.load()
SHAPES = [("stem 422k x 64", 422638, 64, 1), ("L1 256k x 64", 256090, 64, 2), ("L2 69.5k x 128", 69537, 128, 3),
          ("L3 12.7k x 256", 12706, 256, 3), ("L4 2.2k x 512", 2165, 512, 3)]
CFGS = [dict(cr_v4=0), dict(cr_v4=1, cr_cap=4, cr_unroll=4), dict(cr_v4=1, cr_cap=8, cr_unroll=4),
        dict(cr_v4=1, cr_cap=8, cr_unroll=8), dict(cr_v4=1, cr_cap=16, cr_unroll=2), dict(cr_v4=1, cr_cap=8, cr_unroll=2)]
if os.environ.get("PW_CFGS"):
    CFGS = json.loads(os.environ["PW_CFGS"])
NBUF, REPS = 4, 8


def graph_time(fn):
    fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REPS):
            fn(i % NBUF)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3   # microseconds per call


out = []
for cfg in CFGS:
    for k in ("cr_v4", "cr_cap", "cr_unroll"):
        L.set_tuning(k, cfg.get(k, -1))
    row = {"cfg": cfg, "us": {}}
    tot = 0.0
    for name, n, c, calls in SHAPES:
        xs = [torch.randn(n, c, device=dev) for _ in range(NBUF)]
        gs = [torch.randn(n, c, device=dev) for _ in range(NBUF)]
        mean, invstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        ws = torch.empty(2 * c + 1, dtype=torch.float64, device=dev)
        sums = torch.empty(2 * c, device=dev)
        gb = torch.empty(c, device=dev)
        t_stats = graph_time(lambda i: L.call("b2s_bn_stats", xs[i], n, None, c, 1e-5, 0.1, rm, rv, ws, mean, invstd))
        mean.zero_(), invstd.fill_(1.0)
        t_bwd = graph_time(lambda i: L.call("b2s_bn_bwd_reduce", gs[i], xs[i], mean, invstd, gamma, beta, n, None, c, 1,
                                            ws, sums))
        t_col = graph_time(lambda i: L.call("b2s_colsum", gs[i], n, None, c, gb))
        mb = n * c * 4 / 1e6
        row["us"][name] = {"stats": round(t_stats, 1), "bwd_reduce": round(t_bwd, 1), "colsum": round(t_col, 1),
                           "stats_GBps": round(mb / t_stats * 1e3), "bwd_GBps": round(2 * mb / t_bwd * 1e3)}
        tot += calls * (t_stats + t_bwd + t_col)
        del xs, gs
    row["step_us"] = round(tot, 1)
    out.append(row)
    print(json.dumps(row), flush=True)
for k in ("cr_v4", "cr_cap", "cr_unroll"):
    L.set_tuning(k, -1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/pw_bench.json", "w"), indent=1)
