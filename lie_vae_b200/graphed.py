"""CUDA-graph capture of the drop-in autograd surface (host-side plumbing).

At BASELINE configs[1] / configs[2] sizes one forward + backward is 50-100 us of GPU work, while a
``torch.autograd.Function`` round trip (apply, output allocation, the engine's backward thread, ctypes) costs ~150 us of
CPU time -- the public API is launch-latency bound there.  Every kernel launch of this package is
capturable (caller's stream, no allocation or synchronisation inside the C ABI, workspaces come from torch's allocator),
so the same modules run under ``torch.cuda.make_graphed_callables``: forward and backward each replay as ONE graph launch.
This module only packages that recipe for the reference's call pattern (``experiments/vae.py:134-190``):

    hot = graphed_hot_path(model.rep_group, model.decoder, sample_features, n=1)
    x_recon_flat, log_q = hot(features)          # == decoder(group_matrix_to_eazyz(rep_group(features, n))), rep_group.log_posterior()

The graphed callable is bit-identical to the eager modules (tests/test_gpu_graphed.py); noise comes from torch's
generator, which is graph-safe.  Shapes are fixed at capture, as with any CUDA graph.
"""
import torch
from torch import nn

from .lie_tools import group_matrix_to_eazyz

__all__ = ["HotPath", "graphed_hot_path", "graphed"]


class HotPath(nn.Module):
    """encoder features (B, Din) -> (decoder output (n*B, ...), log q (n, B)): the SO(3) latent stretch of ``VAE.forward`` +
    ``VAE.kl`` (``experiments/vae.py:134-146,173-190``) as one module, so that it can be captured as a whole."""

    def __init__(self, rep_group, decoder, n=1):
        super().__init__()
        self.rep_group, self.decoder, self.n = rep_group, decoder, int(n)

    def forward(self, x):
        z = self.rep_group(x, self.n)
        angles = group_matrix_to_eazyz(z.reshape(-1, 3, 3))
        return self.decoder(angles), self.rep_group.log_posterior()


def graphed(fn_or_module, sample_args, num_warmup_iters=3):
    """``torch.cuda.make_graphed_callables`` with this package's conventions (sample tensors on the module's device).

    Garbage is collected first: a CUDA graph (or any cached block) that Python frees *during* a capture makes the allocator
    call ``cudaFree``, which is illegal while a stream is capturing and invalidates the capture (seen when one graphed
    callable is dropped and the next one captured right after)."""
    import gc
    gc.collect()
    torch.cuda.synchronize()
    return torch.cuda.make_graphed_callables(fn_or_module, tuple(sample_args), num_warmup_iters=num_warmup_iters)


def graphed_hot_path(rep_group, decoder, sample_features, n=1):
    """Capture ``HotPath(rep_group, decoder, n)`` for inputs shaped like ``sample_features`` (which must require grad if the
    encoder is to receive gradients)."""
    return graphed(HotPath(rep_group, decoder, n), (sample_features,))
