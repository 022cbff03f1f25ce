"""Micro-benchmark of the batch-norm apply / backward-apply kernels (with GELU, operand twin and, backward, the fused
column sums) at the MSENet14 layer shapes of one 32-plot batch, replayed from a CUDA graph over rotating buffers.
B2S_PW_UNROLL=1 selects the one-row-per-pass kernels.  GPU box only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dpcr_agb_b200 import lib as L

dev = torch.device("cuda:0")
L.load()
SHAPES = [("stem 422k x 64", 422638, 64, 1), ("L1 256k x 64", 256090, 64, 2), ("L2 69.5k x 128", 69537, 128, 3),
          ("L3 12.7k x 256", 12706, 256, 3), ("L4 2.2k x 512", 2165, 512, 3)]
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
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3


tot = 0.0
for name, n, c, calls in SHAPES:
    xs = [torch.randn(n, c, device=dev) for _ in range(NBUF)]
    gs = [torch.randn(n, c, device=dev) for _ in range(NBUF)]
    ys = [torch.empty(n, c, device=dev) for _ in range(NBUF)]
    ts = [torch.empty(n, c, device=dev) for _ in range(NBUF)]
    mean, invstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    sums = torch.zeros(2 * c, device=dev)
    rows = L.query("b2s_bn_bwd_colsum_rows", n, c)
    cs = torch.empty(rows * c, device=dev)
    t_f = graph_time(lambda i: L.call("b2s_bn_apply", xs[i], mean, invstd, gamma, beta, n, None, c, 1, ys[i], ts[i]))
    t_b = graph_time(lambda i: L.call("b2s_bn_bwd_apply", gs[i], xs[i], mean, invstd, gamma, beta, sums, n, None, c, 1, 1,
                                      ys[i], ts[i], cs))
    mb = n * c * 4 / 1e6
    print(f"{name:16s} apply {t_f:6.1f} us {3 * mb / t_f * 1e3:5.0f} GB/s | bwd apply {t_b:6.1f} us {4 * mb / t_b * 1e3:5.0f} GB/s")
    tot += calls * (t_f + t_b)
    del xs, gs, ys, ts
print(f"per step (12 norms): {tot:.1f} us")
