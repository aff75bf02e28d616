#!/usr/bin/env python
"""Run a few micro-batches of the fused pipeline (for ncu / compute-sanitizer captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lie_vae_b200.pipeline import FusedSO3ActionStep  # noqa: E402
import lie_vae_b200.lie_tools as lt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L, C, K = 8, 10, 3
M = (L + 1) ** 2
dev = torch.device("cuda")
torch.manual_seed(0)
mu = lt.random_group_matrices(B, device=dev)
sigma = torch.nn.functional.softplus(torch.randn(B, 3, device=dev))
eps = torch.randn(B, 3, device=dev)
glq = torch.randn(B, device=dev)
item = torch.randn(M, C, device=dev)
gy = torch.randn(B, M * C, device=dev)
y = torch.empty(B, M * C, device=dev)
lq = torch.empty(B, device=dev)
gmu = torch.empty(B, 3, 3, device=dev)
gsg = torch.empty(B, 3, device=dev)
step = FusedSO3ActionStep(B, B, L, C, K, device=dev)
for _ in range(iters):
    step.latent_forward(mu, sigma, eps, lq)
    step.decode_forward(0, B, item, y)
    step.decode_backward(0, B, item, gy)
    step.latent_backward(mu, sigma, eps, glq, gmu, gsg)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()), float(gmu.abs().mean()), float(step.g_item.abs().mean()))
