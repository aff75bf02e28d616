#!/usr/bin/env python
"""Soak test of the warp-decoupled Wigner backward: thousands of launches over many sizes (tile counts per CTA from 0 to
hundreds, ragged tails, both degree specialisations, transpose), each checked bit-for-bit against the first launch of its
configuration and, for the small ones, against the generic cp.async kernel.  Run under `timeout`."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lie_vae_b200 import _ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
t0 = time.time()
launches = 0
sizes = [2, 12, 13, 16, 17, 32, 12 * 147, 12 * 148, 12 * 149 + 1, 12 * 148 * 2, 12 * 148 * 5 + 5, 12 * 148 * 6, 12 * 148 * 7 - 12, 16 * 1000 + 7,
         1 << 16, (1 << 17) + 48, 1 << 18]
for L in (8, 6):
    M = (L + 1) ** 2
    item = torch.randn(M, 10, device=dev)
    for N in sizes:
        for tr in (False, True):
            ang = torch.rand(N, 3, device=dev) * 6 - 3
            g = torch.randn(N, M, 10, device=dev)
            ref = None
            reps = 200 if N <= (1 << 16) else 40
            for i in range(reps):
                a, it = ang.clone().requires_grad_(True), item.clone().requires_grad_(True)
                _ops.WignerApply.apply(a, it, 0, L, tr).backward(g)
                launches += 1
                if ref is None:
                    ref = (a.grad.clone(), it.grad.clone())
                    if N <= 16 * 1000 + 7:       # the generic kernel (unaligned g) as an independent check
                        gu = torch.empty(N * M * 10 + 1, device=dev)[1:].view(N, M, 10)
                        gu.copy_(g)
                        a2, it2 = ang.clone().requires_grad_(True), item.clone().requires_grad_(True)
                        _ops.WignerApply.apply(a2, it2, 0, L, tr).backward(gu)
                        assert (a2.grad - a.grad).abs().max().item() <= 2e-5 * max(1.0, a.grad.abs().max().item()), (L, N, tr)
                        assert (it2.grad - it.grad).abs().max().item() <= 3e-6 * it.grad.abs().max().item(), (L, N, tr)
                else:
                    assert torch.equal(a.grad, ref[0]) and torch.equal(it.grad, ref[1]), (L, N, tr, i)
            torch.cuda.synchronize()
print("soak OK: %d launches in %.1f s" % (launches, time.time() - t0))
