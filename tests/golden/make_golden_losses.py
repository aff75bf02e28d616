"""Generate ``tests/golden/losses.npz`` from the UNMODIFIED reference regularisers (float64).

    python tests/golden/make_golden_losses.py        (build container only: needs /root/reference)

``EquivarianceLoss.forward`` (``losses/equivariance_loss.py:22-48``) draws its angles with ``torch.rand`` and re-encodes the
rotated images through ``model.encode``: here ``torch.rand`` is patched to return stored numbers and the model is a stand-in
whose ``encode`` returns a stored batch of rotation-like matrices, so the fixture pins the SO(3) arithmetic of the real class
(s2s1rodrigues on e_x, bmm, squared distance, mean, lamb) and its autograd gradients.  ``EncoderContinuityLoss`` likewise.
"""
import math
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refshim import load_reference  # noqa: E402

lt, rp, dc = load_reference()
import lie_vae.losses.equivariance_loss as eq  # noqa: E402
import lie_vae.losses.encoder_continuity_loss as ec  # noqa: E402

torch.set_default_dtype(torch.float64)
gen = torch.Generator().manual_seed(20181019)
n = 24
u = torch.rand(n, generator=gen, dtype=torch.float64)                       # what torch.rand returns inside forward
theta = u * 2 * math.pi
enc = lt.quaternions_to_group_matrix(torch.randn(n, 4, generator=gen, dtype=torch.float64)).requires_grad_(True)
enc2 = (lt.quaternions_to_group_matrix(torch.randn(n, 4, generator=gen, dtype=torch.float64))
        + 0.05 * torch.randn(n, 3, 3, generator=gen, dtype=torch.float64)).requires_grad_(True)   # an encoder output need not be exactly orthogonal
img = torch.randn(n, 1, 4, 4, generator=gen, dtype=torch.float64)


class Model:
    def encode(self, x):
        return ((enc2,),)


loss_mod = eq.EquivarianceLoss(Model(), lamb=lambda it: 0.75)
with mock.patch.object(torch, "rand", lambda *a, **k: u.clone()):
    loss = loss_mod(img, enc, 0)
loss.backward()
diffs = loss_mod.diffs[0].detach()

pairs = torch.randn(16, 7, generator=gen, dtype=torch.float64, requires_grad=True)
closs = ec.EncoderContinuityLoss(None, lamb=1.5)(pairs, 0)
closs.backward()

np.savez_compressed(os.path.join(HERE, "losses.npz"), theta=theta.numpy(), enc=enc.detach().numpy(), enc2=enc2.detach().numpy(), lamb=0.75,
                    diffs=diffs.numpy(), loss=loss.detach().numpy(), g_enc=enc.grad.numpy(), g_enc2=enc2.grad.numpy(),
                    pairs=pairs.detach().numpy(), c_lamb=1.5, c_loss=closs.detach().numpy(), g_pairs=pairs.grad.numpy())
print("wrote losses.npz", float(loss), float(closs))
