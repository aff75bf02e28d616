#!/usr/bin/env python
"""Soak test of the persistent double-buffered reparameterize kernels: many launches on the same inputs (plain and
Euler-fused, f32, several sizes with and without a ragged tail) must reproduce the first launch bit for bit -- a race in
the stage hand-over (TMA load into a buffer still being read, store of a buffer being refilled) would show up here."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lie_vae_b200.lie_tools as lt  # noqa: E402
import lie_vae_b200.reparameterize as rp  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
t0, launches = time.time(), 0
for B in ((1 << 20) + 13, 1 << 19, 700 * 256):
    for n in (1, 2):
        if n > 1 and B % 256:
            continue
        mu = lt.random_group_matrices(B, device=dev)
        sg = torch.nn.functional.softplus(torch.randn(B, 3, device=dev))
        eps = torch.randn(n, B, 3, device=dev)
        for fn, w in ((rp.so3_reparameterize, 9), (rp.so3_reparameterize_eazyz, 3)):
            go, gl = torch.randn(n, B, w, device=dev), torch.randn(n, B, device=dev)
            ref = None
            for it in range(150):
                m, s = mu.clone().requires_grad_(True), sg.clone().requires_grad_(True)
                out, lq = fn(m, s, eps, 3)
                torch.autograd.backward([out.reshape(n, B, w), lq], [go, gl])
                launches += 2
                cur = (out.detach(), lq.detach(), m.grad, s.grad)
                if ref is None:
                    ref = cur
                else:
                    for a, b, what in zip(cur, ref, ("out", "log_q", "g_mu", "g_sigma")):
                        if not torch.equal(a, b):
                            raise SystemExit("soak FAILED: %s differs at iteration %d (B=%d n=%d %s)" % (what, it, B, n, fn.__name__))
torch.cuda.synchronize()
print("soak OK: %d launches in %.1f s" % (launches, time.time() - t0))
