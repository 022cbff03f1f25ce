"""Diagnostic (GPU): how far each convolution pass of the tcgen05 kernels is from the EXACT value of its own operand
model (split-bf16 products h*H + l*H + h*L evaluated in fp64) -- i.e. the accumulation error of the tensor-core path
alone -- next to the fp32 SIMT kernels against fp64."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import b2s_testutil as util
from dpcr_agb_b200 import lib
from dpcr_agb_b200.MinkowskiEngine import functional as Fn
from oracle import coords as oc
from oracle import ops as oo

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)


def split(t):
    h, l = oo.split_bf16(t.float())
    return h.double(), l.double()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


for n, cin, cout, K in ((20000, 64, 64, 3), (8000, 128, 128, 3), (3000, 256, 256, 3), (20000, 3, 64, 7)):
    c = util.random_coords(rng, n, nb=2, extent=16 if K == 3 else 12)
    nbr = oc.kernel_map_table(c, c, K, (1, 1, 1))
    x = torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32))
    w = torch.from_numpy((rng.standard_normal((K ** 3, cin, cout)) * 0.05).astype(np.float32))
    gy = torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32))
    nb = oo._as_long(nbr)
    xh, xl = split(x)
    wh, wl = split(w)
    gh, gl = split(gy)
    ex_f = oo._conv_fp32(x.double(), w.double(), nb)                     # exact fp32-operand result
    if cin <= 4:
        w3, _ = oo.split_bf16((w.double() - wh - wl).float())
        m_f = oo._conv_fp32(xh + xl, wh + wl, nb) + oo._conv_fp32(xh, w3.double(), nb)
    else:
        m_f = oo._conv_fp32(xh + xl, wh, nb) + oo._conv_fp32(xh, wl, nb)   # exact value of the split-bf16 model
    wv = torch.zeros_like(w.double()).requires_grad_()
    (ex_w,) = torch.autograd.grad(oo._conv_fp32(x.double(), wv, nb), wv, gy.double())
    if cin <= 4:
        wv = torch.zeros_like(w.double()).requires_grad_()
        (m_w,) = torch.autograd.grad(oo._conv_fp32(xh + xl, wv, nb), wv, gh + gl)
    else:
        wv = torch.zeros_like(w.double()).requires_grad_()
        (a1,) = torch.autograd.grad(oo._conv_fp32(xh + xl, wv, nb), wv, gh)
        wv = torch.zeros_like(w.double()).requires_grad_()
        (a2,) = torch.autograd.grad(oo._conv_fp32(xh, wv, nb), wv, gl)
        m_w = a1 + a2
    xg, wg, gg, ng = x.to(dev), w.to(dev), gy.to(dev), torch.from_numpy(nbr).to(dev)
    lib.set_tuning("precise", 1)
    f_tc = Fn.gather_gemm(xg, wg, None, ng, n, n, cin, cout, K ** 3, 0, impl=2)
    w_tc = Fn.wgrad(xg, gg, ng, n, n, cin, cout, K ** 3, impl=2)
    f_si = Fn.gather_gemm(xg, wg, None, ng, n, n, cin, cout, K ** 3, 0, impl=1)
    w_si = Fn.wgrad(xg, gg, ng, n, n, cin, cout, K ** 3, impl=1)
    print(f"n={n} {cin}->{cout} K={K}: fwd tc vs exact-model {rel(f_tc, m_f):.2e}  model vs fp64 {rel(m_f, ex_f):.2e}  "
          f"simt vs fp64 {rel(f_si, ex_f):.2e} | wgrad tc vs exact-model {rel(w_tc, m_w):.2e}  model vs fp64 "
          f"{rel(m_w, ex_w):.2e}  simt vs fp64 {rel(w_si, ex_w):.2e}", flush=True)
