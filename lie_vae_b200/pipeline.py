"""Pre-allocated, graph-free runner of the fused hot path (host-side plumbing).

One *micro-batch step* is what ``VAE.elbo`` + ``backward`` spend in the SO(3)
latent modules of the reference (``experiments/vae.py:134-190``):

    reparameterize fwd   (mu, sigma, eps)     -> z, log_q        reparameterize.py:220-263
    matrix -> ZYZ Euler   z                   -> angles          vae.py:182, lie_tools.py:178
    Wigner action fwd     angles, item_rep    -> y               decoders.py:47-56
    Wigner action bwd     g_y                 -> g_angles, g_item_rep
    Euler bwd             g_angles            -> g_z
    reparameterize bwd    g_z, g_log_q        -> g_mu, g_sigma

It calls the C ABI directly on caller-owned buffers (no autograd graph, no
allocation in the loop), which is how ``bench.py`` times the kernels with
inputs resident in HBM, and how a training loop that owns its buffers would
drive them.  The autograd modules in ``reparameterize.py`` / ``decoders.py``
are the drop-in API; this class is the same kernels without the tape.
"""
import torch

from . import _cabi
from ._ops import _stream

KERNELS = ("reparam_fwd", "eazyz_fwd", "wigner_fwd", "wigner_bwd", "eazyz_bwd", "reparam_bwd")

# algorithmic HBM bytes per sample of each launch (FP32; SURVEY.md 8d, DESIGN.md "Roofline")
def algorithmic_bytes(L, C):
    M = (L + 1) ** 2
    y = 4 * M * C
    return {
        "reparam_fwd": 36 + 12 + 12 + 36 + 4,          # mu, sigma, eps -> z, log_q
        "eazyz_fwd": 36 + 12,                          # z -> angles
        "wigner_fwd": 12 + y,                          # angles -> y   (item_rep is per step, not per sample)
        "wigner_bwd": y + 12 + 12,                     # g_y, angles -> g_angles
        "eazyz_bwd": 36 + 12 + 36,                     # z, g_angles -> g_z
        "reparam_bwd": 36 + 12 + 12 + 36 + 4 + 36 + 12,  # mu, sigma, eps, g_z, g_lq -> g_mu, g_sigma
    }


class FusedSO3ActionStep:
    """Buffers + launches for a shard of ``shard`` samples (n = 1 sample per datapoint), decoded in
    micro-batches of ``micro`` samples.

    The latent kernels (reparameterize, Euler; ~0.5 kB/sample) run once over the whole shard; only the
    Wigner action, whose output is 3.2 kB/sample, is micro-batched so that y / g_y never need to be
    resident for the full shard.
    """

    LAUNCHES_PER_MICROBATCH = 3      # wigner fwd, wigner bwd (TMA-fed), wigner_reduce_partials
    LAUNCHES_PER_SHARD = 4           # reparam fwd, eazyz fwd, eazyz bwd, reparam bwd

    def __init__(self, shard, micro, degrees=8, rep_copies=10, k=3, transpose=False, device="cuda"):
        self.shard, self.micro = int(shard), int(micro)
        self.L, self.C, self.k, self.transpose = int(degrees), int(rep_copies), int(k), bool(transpose)
        self.M = (self.L + 1) ** 2
        self.device = torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.z = torch.empty((self.shard, 3, 3), **f32)
        self.angles = torch.empty((self.shard, 3), **f32)
        self.g_angles = torch.empty((self.shard, 3), **f32)
        self.g_z = torch.empty((self.shard, 3, 3), **f32)
        self.g_item = torch.empty((self.M, self.C), **f32)      # gradient of the last decoded micro-batch
        with torch.cuda.device(self.device):
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

    def latent_forward(self, mu, sigma, eps, log_q):
        """mu (B,3,3), sigma (B,3), eps (B,3) -> log_q (B); z and the Euler angles stay in the step's buffers."""
        B, st, p = mu.shape[0], _stream(), _cabi.ptr
        self._timed("reparam_fwd", lambda: _cabi.call("lv_so3_reparam_fwd_f32", p(mu), p(sigma), p(eps), p(self.z), p(log_q), 1, B, self.k, st))
        self._timed("eazyz_fwd", lambda: _cabi.call("lv_mat_to_eazyz_fwd_f32", p(self.z), p(self.angles), B, st))

    def decode_forward(self, lo, hi, item_rep, y):
        """Wigner action of samples [lo, hi) on item_rep (M,C) -> y (hi-lo, M, C)."""
        st, p = _stream(), _cabi.ptr
        self._timed("wigner_fwd", lambda: _cabi.call("lv_wigner_apply_fwd_f32", p(self.angles[lo:hi]), p(item_rep), p(y), hi - lo, 0, self.L, self.C, 1, int(self.transpose), st))

    def decode_backward(self, lo, hi, item_rep, g_y):
        """g_y (hi-lo, M, C) -> g_angles[lo:hi] and self.g_item (M,C) for this micro-batch."""
        st, p = _stream(), _cabi.ptr
        self._timed("wigner_bwd", lambda: _cabi.call("lv_wigner_apply_bwd_f32", p(self.angles[lo:hi]), p(item_rep), p(g_y), p(self.g_angles[lo:hi]), p(self.g_item),
                                                    p(self.workspace), self.nws, hi - lo, 0, self.L, self.C, 1, int(self.transpose), st))

    def latent_backward(self, mu, sigma, eps, g_log_q, g_mu, g_sigma):
        """g_angles (from decode_backward) and g_log_q (B) -> g_mu (B,3,3), g_sigma (B,3)."""
        B, st, p = mu.shape[0], _stream(), _cabi.ptr
        self._timed("eazyz_bwd", lambda: _cabi.call("lv_mat_to_eazyz_bwd_f32", p(self.z), p(self.g_angles), p(self.g_z), B, st))
        self._timed("reparam_bwd", lambda: _cabi.call("lv_so3_reparam_bwd_f32", p(mu), p(sigma), p(eps), p(self.g_z), p(g_log_q), p(g_mu), p(g_sigma), 1, B, self.k, st))
