"""Pre-allocated, graph-free runner of the fused hot path (host-side plumbing).

One *step* is what ``VAE.elbo`` + ``backward`` spend in the SO(3) latent modules of the reference
(``experiments/vae.py:134-190``), as four kernels:

    latent fwd    (mu, sigma, eps)          -> angles, log_q     reparameterize.py:220-263 + vae.py:182 / lie_tools.py:178
    Wigner fwd    angles, item_rep          -> y                 decoders.py:47-56
    Wigner bwd    g_y                       -> g_angles, g_item_rep
    latent bwd    g_angles, g_log_q         -> g_mu, g_sigma

The latent kernels are the reparameterize kernels fused with matrix -> ZYZ Euler: the sampled pose z never
touches HBM (the backward recomputes it from mu, sigma, eps).  This class calls the C ABI directly on
caller-owned buffers (no autograd graph, no allocation in the loop), which is how ``bench.py`` times the
kernels with inputs resident in HBM, and how a training loop that owns its buffers would drive them.  The
autograd modules in ``reparameterize.py`` / ``decoders.py`` are the drop-in API; this is the same kernels
without the tape.
"""
import torch

from . import _cabi
from ._ops import _on, _stream

KERNELS = ("latent_fwd", "wigner_fwd", "wigner_bwd", "latent_bwd")


def algorithmic_bytes(L, C, in_kernel_noise=False):
    """Algorithmic HBM bytes per sample of each launch (FP32).  With in-kernel noise
    eps (12 B) is neither read by the forward nor by the backward."""
    M = (L + 1) ** 2
    y = 4 * M * C
    e = 0 if in_kernel_noise else 12
    return {
        "latent_fwd": 36 + 12 + e + 12 + 4,             # mu, sigma, eps -> angles, log_q
        "wigner_fwd": 12 + y,                           # angles -> y   (item_rep is per step, not per sample)
        "wigner_bwd": y + 12 + 12,                      # g_y, angles -> g_angles
        "latent_bwd": 36 + 12 + e + 12 + 4 + 36 + 12,   # mu, sigma, eps, g_angles, g_lq -> g_mu, g_sigma
    }


class FusedSO3ActionStep:
    """Buffers + launches for a shard of ``shard`` samples (n = 1 sample per datapoint), decoded in
    micro-batches of ``micro`` samples.

    The latent kernels (~0.2 kB/sample) run once over the whole shard; only the Wigner action, whose output
    is 3.2 kB/sample, is micro-batched so that y / g_y never need to be resident for the full shard.
    """

    LAUNCHES_PER_MICROBATCH = 3      # wigner fwd, wigner bwd (degree-specialised, TMA-fed), wigner_reduce_partials
    LAUNCHES_PER_SHARD = 2           # latent fwd, latent bwd

    def __init__(self, shard, micro, degrees=8, rep_copies=10, k=3, transpose=False, device="cuda"):
        self.shard, self.micro = int(shard), int(micro)
        self.L, self.C, self.k, self.transpose = int(degrees), int(rep_copies), int(k), bool(transpose)
        self.M = (self.L + 1) ** 2
        self.device = torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.angles = torch.empty((self.shard, 3), **f32)
        self.g_angles = torch.empty((self.shard, 3), **f32)
        self.g_item = torch.empty((self.M, self.C), **f32)      # gradient of the last decoded micro-batch
        with _on(self.device):
            nws = _cabi.lib().lv_wigner_bwd_workspace_floats(self.micro, 0, self.L, self.C)
        if nws < 0:
            raise RuntimeError(_cabi.last_error())
        self.nws = nws
        self.workspace = torch.empty(max(nws, 1), **f32)
        self.events = None       # optional per-kernel CUDA events: {name: [(start, stop), ...]}

    def enable_kernel_timing(self, on=True):
        self.events = {k: [] for k in KERNELS} if on else None

    def _timed(self, name, fn):
        if self.events is None:
            fn()
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        self.events[name].append((a, b))

    def latent_forward(self, mu, sigma, eps, log_q, z=None, seed=0, offset=0):
        """mu (B,3,3), sigma (B,3), eps (B,3) -> log_q (B); the Euler angles of the sampled pose stay in the
        step's buffer, the pose itself is written only if ``z`` (B,3,3) is given.  ``eps=None``: in-kernel noise
        (Philox keyed by ``seed``, counter ``offset`` + sample index; pass the same pair to ``latent_backward``)."""
        B, p = mu.shape[0], _cabi.ptr
        with _on(self.device):
            st = _stream()
            if eps is None:
                self._timed("latent_fwd", lambda: _cabi.call("lv_so3_reparam_philox_fwd_f32", p(mu), p(sigma), int(seed), int(offset), p(z),
                                                            p(self.angles), p(log_q), 1, B, self.k, st))
            else:
                self._timed("latent_fwd", lambda: _cabi.call("lv_so3_reparam_eazyz_fwd_f32", p(mu), p(sigma), p(eps), p(z), p(self.angles),
                                                            p(log_q), 1, B, self.k, st))

    def decode_forward(self, lo, hi, item_rep, y):
        """Wigner action of samples [lo, hi) on item_rep (M,C) -> y (hi-lo, M, C)."""
        p = _cabi.ptr
        with _on(self.device):
            st = _stream()
            self._timed("wigner_fwd", lambda: _cabi.call("lv_wigner_apply_fwd_f32", p(self.angles[lo:hi]), p(item_rep), p(y), hi - lo, 0, self.L,
                                                        self.C, 1, int(self.transpose), st))

    def decode_backward(self, lo, hi, item_rep, g_y, accumulate=False):
        """g_y (hi-lo, M, C) -> g_angles[lo:hi] and self.g_item (M,C): the gradient of this micro-batch, or, with
        ``accumulate``, added to what self.g_item already holds (zero it at the start of a step)."""
        p = _cabi.ptr
        with _on(self.device):
            st = _stream()
            self._timed("wigner_bwd", lambda: _cabi.call("lv_wigner_apply_bwd_f32", p(self.angles[lo:hi]), p(item_rep), p(g_y), p(self.g_angles[lo:hi]),
                                                        p(self.g_item), p(self.workspace), self.nws, hi - lo, 0, self.L, self.C,
                                                        3 if accumulate else 1, int(self.transpose), st))

    def latent_backward(self, mu, sigma, eps, g_log_q, g_mu, g_sigma, g_z=None, seed=0, offset=0):
        """g_angles (from decode_backward), g_log_q (B) and optionally g_z (B,3,3) -> g_mu (B,3,3), g_sigma (B,3)."""
        B, p = mu.shape[0], _cabi.ptr
        with _on(self.device):
            st = _stream()
            if eps is None:
                self._timed("latent_bwd", lambda: _cabi.call("lv_so3_reparam_philox_bwd_f32", p(mu), p(sigma), int(seed), int(offset), p(g_z),
                                                            p(self.g_angles), p(g_log_q), p(g_mu), p(g_sigma), 1, B, self.k, st))
            else:
                self._timed("latent_bwd", lambda: _cabi.call("lv_so3_reparam_eazyz_bwd_f32", p(mu), p(sigma), p(eps), p(g_z), p(self.g_angles),
                                                            p(g_log_q), p(g_mu), p(g_sigma), 1, B, self.k, st))
