"""``lie_vae.utils`` mirror."""
import torch


def logsumexp(inputs, dim=None, keepdim=False):
    """Numerically stable log-sum-exp   (``utils.py:4-26``).

    The wrapped-density use (``reparameterize.py:261``) is fused into the SO(3)
    reparameterize kernel; this stand-alone entry serves ``vae.py:171`` (IWAE
    over the sample axis) and delegates to ATen's fused ``torch.logsumexp``.
    """
    if dim is None:
        inputs = inputs.reshape(-1)
        dim = 0
    return torch.logsumexp(inputs, dim=dim, keepdim=keepdim)
