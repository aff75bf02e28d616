"""Encoder regularisers -- drop-in for the reference's ``lie_vae.losses``.

``EquivarianceLoss`` (``losses/equivariance_loss.py:11-57``) asks the encoder to commute with in-plane image rotations:
for a random angle theta the encoding of the rotated image should be ``Rx(theta) . encoding``.  Its SO(3) part -- building
``Rx = s2s1rodrigues(e_x, (cos, sin))``, the batched product and the squared distance -- is one sm_100a kernel per direction
(``_ops.EquivarianceSqDist``); the image warp is ``torch.nn.functional`` as in the reference.  ``EncoderContinuityLoss``
(``losses/encoder_continuity_loss.py:6-38``) is a pairwise squared distance of encodings: plain tensor arithmetic on whatever
device the encodings live on.  Constructor arguments, ``forward`` signatures, the ``diffs`` buffer and the logging cadence are
the reference's.
"""
from math import pi

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _ops

__all__ = ["EquivarianceLoss", "EncoderContinuityLoss", "equivariance_sqdist"]


def equivariance_sqdist(theta, encoding, encoding_of_rotated):
    """(n) squared distances || Rx(theta) encoding - encoding_of_rotated ||_F^2 (CUDA float32 / float64; differentiable in
    both encodings)."""
    return _ops.EquivarianceSqDist.apply(theta, encoding, encoding_of_rotated)


def _weight(lamb, it):
    return lamb(it) if callable(lamb) else lamb


class _Regulariser(nn.Module):
    tag = ""

    def __init__(self, model, lamb=1.0, log=None, report_freq=1):
        super().__init__()
        self.model, self.lamb, self.log, self.report_freq = model, lamb, log, report_freq
        self.diffs = []

    def _finish(self, diffs, it):
        """mean(diffs) * lamb(it); every ``report_freq`` iterations the running mean and the weight go to the logger."""
        self.diffs.append(diffs)
        lamb = _weight(self.lamb, it)
        if self.log and (it + 1) % self.report_freq == 0:
            self.log.add_scalar(self.tag, torch.cat(self.diffs).mean(), it + 1)
            self.log.add_scalar(self.tag + "_lamb", lamb, it + 1)
            self.diffs = []
        return diffs.mean() * lamb


class EquivarianceLoss(_Regulariser):
    """Equivariance of the encoder under the SO(2) subgroup of in-plane rotations (``equivariance_loss.py:11-57``).  The
    reference requires a callable ``lamb``; a constant is accepted too."""
    tag = "equivariance"

    def __init__(self, model, num_samples=None, lamb=1.0, log=None, report_freq=1):
        super().__init__(model, lamb, log, report_freq)
        self.num_samples = num_samples

    def forward(self, img, encoding, it):
        assert encoding.shape[-2:] == (3, 3), "Rotation matrix input required"
        if self.num_samples:
            img, encoding = img[:self.num_samples], encoding[:self.num_samples]
        theta = torch.rand(img.shape[0], device=encoding.device) * 2 * pi
        rotated_encoding = self.model.encode(self.rotate(img, theta))[0][0]
        return self._finish(equivariance_sqdist(theta.to(encoding.dtype), encoding, rotated_encoding), it)

    @staticmethod
    def rotate(img, theta):
        """The batch of images turned by theta about their centres (bilinear resampling, as ``equivariance_loss.py:50-57``)."""
        c, s, o = torch.cos(theta), torch.sin(theta), torch.zeros_like(theta)
        grid = F.affine_grid(torch.stack([c, -s, o, s, c, o], 1).view(-1, 2, 3), img.size(), align_corners=False)
        return F.grid_sample(img, grid, align_corners=False)


class EncoderContinuityLoss(_Regulariser):
    """Squared distance between the encodings of consecutive input pairs (rows 2i and 2i+1;
    ``encoder_continuity_loss.py:6-38``)."""
    tag = "encoder_continuity"

    def forward(self, encodings, it):
        pairs = encodings.reshape(encodings.shape[0] // 2, 2, -1)
        return self._finish((pairs[:, 0] - pairs[:, 1]).pow(2).sum(-1), it)
