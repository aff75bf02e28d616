#!/usr/bin/env python
"""Small Wigner / reparameterize launches for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family
once, sizes chosen so the warp-decoupled backward runs several tiles per CTA on a handful of CTAs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lie_vae_b200.lie_tools as lt  # noqa: E402
import lie_vae_b200.reparameterize as rp  # noqa: E402
from lie_vae_b200 import _ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
for N, L, C, shared in ((16 * 40 + 5, 8, 10, True), (16 * 9, 6, 10, True), (333, 8, 10, False), (100, 3, 3, True), (64, 11, 2, True)):
    M = (L + 1) ** 2
    ang = (torch.rand(N, 3, device=dev) * 6 - 3).requires_grad_(True)
    spec = (torch.randn(M, C, device=dev) if shared else torch.randn(N, M, C, device=dev)).requires_grad_(True)
    out = _ops.wigner_apply(ang, spec, 0, L, False)
    out.backward(torch.randn_like(out))
    torch.cuda.synchronize()
    print("wigner", N, L, C, shared, float(out.abs().mean()), float(spec.grad.abs().mean()), float(ang.grad.abs().mean()))
mu = lt.random_group_matrices(1000, device=dev).requires_grad_(True)
sg = torch.rand(1000, 3, device=dev).add(0.1).requires_grad_(True)
for fn in (rp.so3_reparameterize, rp.so3_reparameterize_eazyz):
    a, lq = fn(mu, sg, torch.randn(2, 1000, 3, device=dev), 3)
    (a.sum() + lq.sum()).backward()
torch.cuda.synchronize()
print("ok", float(mu.grad.abs().mean()))
