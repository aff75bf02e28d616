"""``lie_vae.utils`` mirror."""
import math

import torch

from . import _ops

__all__ = ["logsumexp", "iwae_log_likelihood", "action_log_likelihood"]


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


def action_log_likelihood(rep_group, decoder, features, x, n=500):
    """``VAE.log_likelihood`` (``experiments/vae.py:164-171``; called with batch 1 and n = 500, ``main.py:134-143``) for an
    SO(3) latent with an ``ActionNet`` decoder whose ``deconv`` is the identity (the toy configuration): n importance samples
    per datapoint, reconstruction term fused into the Wigner forward (``_ops.wigner_recon_sse``: the (n*B, M, C) decoder
    output is never written) and the log-sum-exp over n in one more kernel.  ``features`` (B, Din) = encoder output,
    ``x`` (B, M, C) the data.  Evaluation only: no gradients."""
    from .lie_tools import group_matrix_to_eazyz
    with torch.no_grad():
        z = rep_group(features, n)                                        # (n, B, 3, 3)
        B = z.shape[1]
        angles = group_matrix_to_eazyz(z.reshape(-1, 3, 3))
        sse = _ops.wigner_recon_sse(angles, decoder.item_rep, x.reshape(B, decoder.matrix_dims, decoder.rep_copies),
                                    decoder.degrees, decoder.transpose).view(n, B)
        return iwae_log_likelihood(-sse, rep_group.log_prior(), rep_group.log_posterior())
