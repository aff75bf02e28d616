#!/usr/bin/env python
"""Where the ~0.1 ms of host time per autograd fwd+bwd pair goes (BASELINE configs[1] size; CPU-bound loop, wall clock)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lie_vae_b200.lie_tools as lt  # noqa: E402
import lie_vae_b200.reparameterize as rp  # noqa: E402
from lie_vae_b200 import _cabi  # noqa: E402
from lie_vae_b200._ops import _stream  # noqa: E402

dev = torch.device("cuda")
B = 1 << 16          # small kernels: the loop is host-bound
mu = lt.random_group_matrices(B, device=dev).requires_grad_(True)
sg = torch.nn.functional.softplus(torch.randn(B, 3, device=dev)).requires_grad_(True)
eps, gz, glq = torch.randn(1, B, 3, device=dev), torch.randn(1, B, 3, 3, device=dev), torch.randn(1, B, device=dev)
N = 2000


def wall(fn):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(N):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / N * 1e6


def fwd_nograd():
    with torch.no_grad():
        rp.so3_reparameterize(mu, sg, eps, 3)


def fwd_grad():
    rp.so3_reparameterize(mu, sg, eps, 3)


def fwd_bwd():
    mu.grad = sg.grad = None
    z, lq = rp.so3_reparameterize(mu, sg, eps, 3)
    torch.autograd.backward([z, lq], [gz, glq])


zb, lqb = torch.empty(1, B, 3, 3, device=dev), torch.empty(1, B, device=dev)
p_ = _cabi.ptr


def cabi_fwd():
    _cabi.call("lv_so3_reparam_fwd_f32", p_(mu), p_(sg), p_(eps), p_(zb), p_(lqb), 1, B, 3, _stream())


def torch_pair():          # two trivial torch ops through autograd, for scale
    mu.grad = None
    (mu * 2.0).backward(gz[0])


for name, fn in (("C ABI forward call (ctypes)", cabi_fwd), ("forward, no_grad", fwd_nograd), ("forward, grad mode", fwd_grad),
                 ("forward + backward", fwd_bwd), ("torch mul + backward (scale)", torch_pair)):
    print("%-34s %7.1f us" % (name, wall(fn)))
pr = cProfile.Profile()
pr.enable()
for _ in range(1000):
    fwd_bwd()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
