"""``lie_vae.utils`` mirror."""
import math

import torch

from . import _ops

__all__ = ["logsumexp", "iwae_log_likelihood"]


def logsumexp(inputs, dim=None, keepdim=False):
    """Numerically stable log-sum-exp   (``utils.py:4-26``), one sm_100a kernel (single pass, running max and sum).

    The wrapped-density use (``reparameterize.py:261``) is fused into the SO(3) reparameterize kernel; this stand-alone
    entry serves ``vae.py:171`` (IWAE over the sample axis).  CUDA float32 / float64 tensors.
    """
    if dim is None:
        inputs = inputs.reshape(-1)
        dim = 0
    dim = dim % inputs.dim()
    out = _ops.LogSumExpLeading.apply(inputs.movedim(dim, 0))
    return out.unsqueeze(dim) if keepdim else out


def iwae_log_likelihood(log_p_x_z, log_p_z, log_q_z_x):
    """Importance-weighted bound of ``VAE.log_likelihood`` (``experiments/vae.py:164-171``):
    ``(logsumexp_n(log p(x|z) + log p(z) - log q(z|x)) - log n).mean()`` for (n,B) inputs."""
    w = log_p_x_z + log_p_z - log_q_z_x
    return (logsumexp(w, dim=0) - math.log(w.shape[0])).mean()
