"""SO(3) reparameterization modules -- drop-in for the reference's ``lie_vae.reparameterize``.

Same class names, constructor arguments, attributes and ``state_dict`` keys
(``reparameterize.py:100-278``).  ``SO3reparameterize.forward`` runs ONE fused
sm_100a kernel that scales the algebra noise, exponentiates it (Rodrigues),
left-multiplies by the mean rotation and evaluates the wrapped log-density with
its 2k+1 winding terms; ``log_posterior()`` returns that kernel's second output.
The Euclidean (``Nreparameterize``) and vMF (``Sreparameterize``) baselines of the
reference (``reparameterize.py:16-97``) are not on the SO(3) hot path; they are kept as
plain PyTorch modules with the reference's names, attributes and ``state_dict`` keys so that
``experiments/vae.py:7-9`` imports resolve against this package unchanged.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _ops
from .lie_tools import rodrigues, quaternions_to_group_matrix, s2s1rodrigues, s2s2_gram_schmidt

__all__ = ["Nreparameterize", "Sreparameterize", "N0reparameterize", "AlgebraMean", "QuaternionMean", "S2S1Mean", "S2S2Mean", "SO3reparameterize",
           "so3_reparameterize", "so3_reparameterize_eazyz", "so3_head_reparameterize", "LOG_PRIOR_SO3"]

LOG_PRIOR_SO3 = -math.log(8.0 * math.pi ** 2)   # reparameterize.py:266


def so3_reparameterize(mu, sigma, eps, k=10):
    """Functional form of the fused kernel.

    mu (B,3,3) mean rotations, sigma (B,3) algebra scales, eps (n,B,3) standard-normal noise.
    Returns z = mu @ exp(hat(eps*sigma)) of shape (n,B,3,3) and log q(z|x) of shape (n,B).
    Differentiable in mu and sigma.
    """
    return _ops.so3_reparam(mu, sigma, eps, k, False)


def so3_reparameterize_eazyz(mu, sigma, eps, k=10):
    """``group_matrix_to_eazyz(so3_reparameterize(...)[0])`` and the log-density in ONE kernel.

    Returns the ZYZ Euler angles (n,B,3) of the sampled pose -- what ``VAE.decode`` passes to the action
    decoder (``experiments/vae.py:182``) -- and log q (n,B); the 3x3 pose never leaves registers.
    Differentiable in mu and sigma.
    """
    return _ops.so3_reparam(mu, sigma, eps, k, True)


def so3_reparameterize_philox(mu, sigma, n=1, k=10, seed=0, offset=0, euler=False):
    """The fused kernel with IN-KERNEL noise: eps ~ N(0,1) of ``reparameterize.py:137-141`` is generated per sample by
    Philox4x32-10 (key ``seed``, counter ``offset`` + flat sample index) and regenerated in the backward, so it never
    exists in memory.  Returns (z (n,B,3,3) -- or its ZYZ Euler angles (n,B,3) with ``euler`` --, log_q (n,B)).
    ``philox_normal(n * B, seed, offset).view(n, B, 3)`` is the eps it uses (same bits)."""
    return _ops.SO3ReparamPhilox.apply(mu, sigma, n, k, seed, offset, euler)


philox_normal = _ops.philox_normal


_HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


def _normal_log_prob(z, mean, std):
    return -((z - mean) ** 2) / (2 * std ** 2) - std.log() - _HALF_LOG_2PI


class Nreparameterize(nn.Module):
    """Diagonal Gaussian latent in R^z_dim (``reparameterize.py:16-55``): the reference's Euclidean baseline.
    Host-side PyTorch (two Linear heads + softplus); nothing here touches the SO(3) kernels."""

    def __init__(self, input_dim, z_dim):
        super().__init__()
        self.input_dim = input_dim
        self.z_dim = z_dim
        self.sigma_linear = nn.Linear(input_dim, z_dim)
        self.mu_linear = nn.Linear(input_dim, z_dim)
        self.return_means = False
        self.mu, self.sigma, self.z = None, None, None

    def forward(self, x, n=1):
        self.mu = self.mu_linear(x)
        self.sigma = F.softplus(self.sigma_linear(x))
        self.z = self.nsample(n=n)
        return self.z

    def kl(self):
        return -0.5 * torch.sum(1 + 2 * self.sigma.log() - self.mu.pow(2) - self.sigma ** 2, -1)

    def log_posterior(self):
        return self._log_posterior(self.z)

    def _log_posterior(self, z):
        return _normal_log_prob(z, self.mu, self.sigma).sum(-1)

    def log_prior(self):
        return _normal_log_prob(self.z, torch.zeros_like(self.mu), torch.ones_like(self.sigma)).sum(-1)

    def nsample(self, n=1):
        if self.return_means:
            return self.mu.expand(n, -1, -1)
        eps = torch.randn((n,) + tuple(self.mu.shape), dtype=self.mu.dtype, device=self.mu.device)
        return self.mu + eps * self.sigma

    def deterministic(self):
        """Set to return means."""
        self.return_means = True


class Sreparameterize(nn.Module):
    """von Mises-Fisher latent on the sphere (``reparameterize.py:58-97``).  The vMF arithmetic lives in the
    third-party ``hyperspherical_vae_pytorch`` package (``reparameterize.py:13``), which this package does not
    restate: the module keeps the reference's surface and raises ``ImportError`` at first use if it is absent."""

    def __init__(self, input_dim, z_dim):
        super().__init__()
        self.input_dim = input_dim
        self.z_dim = z_dim
        self.k_linear = nn.Linear(input_dim, 1)
        self.mu_linear = nn.Linear(input_dim, z_dim)
        self.return_means = False
        self.mu, self.k, self.z = None, None, None

    @staticmethod
    def _dists():
        try:
            from hyperspherical_vae_pytorch.distributions import VonMisesFisher, HypersphericalUniform
        except ImportError as e:
            raise ImportError("Sreparameterize needs the hyperspherical_vae_pytorch package (vMF sampling / entropy), "
                              "as the reference does (reparameterize.py:13)") from e
        return VonMisesFisher, HypersphericalUniform

    def forward(self, x, n=1):
        mu = self.mu_linear(x)
        self.mu = mu / mu.norm(p=2, dim=-1, keepdim=True)
        self.k = F.softplus(self.k_linear(x)) + 1
        self.z = self.nsample(n=n)
        return self.z

    def kl(self):
        vmf, unif = self._dists()
        return -vmf(self.mu, self.k).entropy() + unif(self.z_dim - 1).entropy().to(self.mu.device)

    def log_posterior(self):
        return self._dists()[0](self.mu, self.k).log_prob(self.z)

    def log_prior(self):
        return self._dists()[1](self.z_dim - 1).log_prob(self.z)

    def nsample(self, n=1):
        if self.return_means:
            return self.mu.expand(n, -1, -1)
        return self._dists()[0](self.mu, self.k).rsample(n)

    def deterministic(self):
        """Set to return means."""
        self.return_means = True


def so3_head_reparameterize(h, mean_weight, mean_bias, sigma_weight, sigma_bias, eps, mode, k=10, euler=False):
    """Encoder heads + reparameterize in ONE kernel.

    h (B,Din <= 32) encoder features; ``mean_weight`` (Dm,Din) / ``mean_bias`` (Dm) the mean head's Linear with
    mode 'alg' (Dm = 3, ``AlgebraMean``), 'q' (Dm = 4, ``QuaternionMean``), 's2s2' (Dm = 6, ``S2S2Mean``) or 's2s1'
    (Dm = 5, ``S2S1Mean``: rows of ``s2_map`` then ``s1_map``);
    ``sigma_weight`` (3,Din) / ``sigma_bias`` (3) the sigma head's; eps (n,B,3).
    Returns (z (n,B,3,3) -- or its ZYZ Euler angles (n,B,3) with ``euler`` --, log_q (n,B), mu (B,3,3), sigma (B,3)).
    Differentiable in h and the four head parameters.
    """
    return _ops.SO3HeadReparam.apply(h, mean_weight, mean_bias, sigma_weight, sigma_bias, eps, mode, k, euler)


class N0reparameterize(nn.Module):
    """Zero-mean Gaussian in the algebra (``reparameterize.py:100-145``)."""

    def __init__(self, input_dim, z_dim, fixed_sigma=None):
        super().__init__()
        self.input_dim = input_dim
        self.z_dim = z_dim
        self.sigma_linear = nn.Linear(input_dim, z_dim)
        self.return_means = False
        if fixed_sigma is not None:
            self.register_buffer('fixed_sigma', torch.tensor(fixed_sigma))
        else:
            self.fixed_sigma = None
        self.sigma = None
        self.z = None
        self.eps = None

    def compute_sigma(self, x):
        if self.fixed_sigma is not None:
            return x.new_full((x.shape[0], self.z_dim), float(self.fixed_sigma))
        return F.softplus(self.sigma_linear(x))

    def sample_noise(self, n=1, like=None):
        """Standard-normal noise (n,B,z_dim) for the current sigma (or for the batch of ``like`` (B, ...) before sigma
        exists: the fused-head path samples first); zeros when deterministic."""
        ref = self.sigma if like is None else like
        shape = (n, ref.shape[0], self.z_dim)
        if self.return_means:
            return ref.new_zeros(shape)
        return torch.randn(shape, dtype=ref.dtype, device=ref.device)

    def forward(self, x, n=1):
        self.sigma = self.compute_sigma(x)
        self.z = self.nsample(n=n)
        return self.z

    def kl(self):
        return -0.5 * torch.sum(1 + 2 * self.sigma.log() - self.sigma ** 2, -1)

    def log_posterior(self):
        return self._log_posterior(self.z)

    def _log_posterior(self, z):
        s = self.sigma
        return (-(z ** 2) / (2 * s ** 2) - s.log() - 0.5 * math.log(2 * math.pi)).sum(-1)

    def log_prior(self):
        return (-(self.z ** 2) / 2 - 0.5 * math.log(2 * math.pi)).sum(-1)

    def nsample(self, n=1):
        self.eps = self.sample_noise(n)
        if self.return_means:
            return torch.zeros_like(self.sigma).expand(n, -1, -1)
        return self.eps * self.sigma

    def deterministic(self):
        """Set to return means."""
        self.return_means = True


class AlgebraMean(nn.Module):
    """R^3 -> SO(3) through the exponential map (``reparameterize.py:148-155``)."""

    def __init__(self, input_dims):
        super().__init__()
        self.map = nn.Linear(input_dims, 3)

    def forward(self, x):
        return rodrigues(self.map(x))


class QuaternionMean(nn.Module):
    """R^4 -> SO(3) through normalised quaternions (``reparameterize.py:158-164``)."""

    def __init__(self, input_dims):
        super().__init__()
        self.map = nn.Linear(input_dims, 4)

    def forward(self, x):
        return quaternions_to_group_matrix(self.map(x))


class S2S1Mean(nn.Module):
    """R^5 -> SO(3): unit axis and unit (cos, sin) (``reparameterize.py:167-181``)."""

    def __init__(self, input_dims):
        super().__init__()
        self.s2_map = nn.Linear(input_dims, 3)
        self.s1_map = nn.Linear(input_dims, 2)

    def forward(self, x):
        s2_el = self.s2_map(x)
        s2_el = s2_el / s2_el.norm(p=2, dim=-1, keepdim=True)
        s1_el = self.s1_map(x)
        s1_el = s1_el / s1_el.norm(p=2, dim=-1, keepdim=True)
        return s2s1rodrigues(s2_el, s1_el)


class S2S2Mean(nn.Module):
    """R^6 -> SO(3) by Gram-Schmidt, evaluated in float64 (``reparameterize.py:184-197``)."""

    def __init__(self, input_dims):
        super().__init__()
        self.map = nn.Linear(input_dims, 6)
        # Start with big outputs
        self.map.weight.data.uniform_(-10, 10)
        self.map.bias.data.uniform_(-10, 10)

    def forward(self, x):
        v = self.map(x).double().view(-1, 2, 3)
        v1, v2 = v[:, 0], v[:, 1]
        return s2s2_gram_schmidt(v1, v2).float()


class SO3reparameterize(nn.Module):
    """Reparameterized SO(3) latent (``reparameterize.py:200-278``).

    ``forward(x, n)`` -> z (n,B,3,3).  mu = mean_module(x), sigma from the inner
    ``N0reparameterize``; sampling, exp-map, composition and the wrapped log-density
    run in one kernel.  ``log_posterior()`` -> (n,B) float32, ``log_prior()`` -> (n,B)
    float64 constant, ``kl()`` -> (B,) float64, exactly the reference's dtypes.
    """

    def __init__(self, reparameterize, mean_module, k=10):
        super().__init__()
        self.mean_module = mean_module
        self.reparameterize = reparameterize
        self.input_dim = self.reparameterize.input_dim
        assert self.reparameterize.z_dim == 3
        self.k = k
        self.return_means = False
        self._v, self._philox = None, None
        self.mu_lie, self.v, self.z = None, None, None
        self._log_q = None
        # in_kernel_noise: eps is generated inside the kernel (Philox keyed by torch's seed, counter = samples drawn so far)
        # instead of being drawn by torch and stored -- the throughput mode of SURVEY.md section 7; the unfused path only
        self.in_kernel_noise = False
        self._noise_offset = 0
        # encoder heads (Linear + mean map, Linear + softplus) inside the reparameterize kernel when the mean module
        # is one of the four reference mean maps on float32 CUDA features; set False for the unfused launches
        self.fuse_heads = True
        self._fused = False                                 # whether the last forward ran the fused-head kernel

    @property
    def v(self):
        """Algebra sample eps * sigma (n,B,3) (``reparameterize.py:222``); with in-kernel noise it is materialised on demand."""
        if self._v is None and self._philox is not None:
            seed, offset, n, B = self._philox
            rep = self.reparameterize
            rep.eps = _ops.philox_normal(n * B, seed, offset, dtype=rep.sigma.dtype, device=rep.sigma.device).view(n, B, 3)
            self._v = rep.eps * rep.sigma
            rep.z = self._v
        return self._v

    @v.setter
    def v(self, value):
        self._v = value
        self._philox = None

    def _fused_head_mode(self, x):
        """'alg' / 'q' / 's2s2' when the encoder heads can run inside the reparameterize kernel, else None."""
        mode = {AlgebraMean: "alg", QuaternionMean: "q", S2S2Mean: "s2s2", S2S1Mean: "s2s1"}.get(type(self.mean_module))
        rep = self.reparameterize
        ok = (self.fuse_heads and not self.in_kernel_noise and mode is not None and type(rep) is N0reparameterize and rep.fixed_sigma is None
              and not self.return_means and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
              and x.shape[1] <= _ops.HEAD_MAX_DIN and rep.sigma_linear.weight.dtype == torch.float32)
        return mode if ok else None

    def _forward_fused(self, x, n, mode):
        rep, mean = self.reparameterize, self.mean_module
        rep.sigma = None
        try:
            rep.eps = rep.sample_noise(n, like=x)
        except TypeError:                                   # a user-supplied sample_noise(n) without the keyword
            rep.eps = rep.sample_noise(n)
        if mode == "s2s1":      # two Linear layers feed this mean map: stack them (axis rows, then (cos, sin) rows)
            wm, bm = torch.cat([mean.s2_map.weight, mean.s1_map.weight], 0), torch.cat([mean.s2_map.bias, mean.s1_map.bias], 0)
        else:
            wm, bm = mean.map.weight, mean.map.bias
        self.z, self._log_q, self.mu_lie, rep.sigma = _ops.SO3HeadReparam.apply(
            x, wm, bm, rep.sigma_linear.weight, rep.sigma_linear.bias, rep.eps, mode, self.k, False)
        self.v = rep.eps * rep.sigma
        rep.z = self.v
        self._fused = True
        return self.z

    def forward(self, x, n=1):
        mode = self._fused_head_mode(x)
        if mode is not None:
            return self._forward_fused(x, n, mode)
        self._fused = False
        self.mu_lie = self.mean_module(x)
        rep = self.reparameterize
        rep.sigma = rep.compute_sigma(x)
        if self.in_kernel_noise and not self.return_means and type(rep) is N0reparameterize:
            B = rep.sigma.shape[0]
            seed, offset = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, self._noise_offset
            self._noise_offset += n * B
            self.z, self._log_q = so3_reparameterize_philox(self.mu_lie, rep.sigma, n, self.k, seed, offset)
            self._v, self._philox = None, (seed, offset, n, B)      # v / eps are materialised only if somebody reads them
            rep.eps, rep.z = None, None
            return self.z
        rep.eps = rep.sample_noise(n)
        self.z, self._log_q = so3_reparameterize(self.mu_lie, rep.sigma, rep.eps, self.k)
        self.v = rep.eps * rep.sigma          # attribute parity; not consumed by the kernel path
        rep.z = self.v
        if self.return_means:
            self.z = self.mu_lie.expand(n, *[-1] * len(self.mu_lie.shape))
        return self.z

    def nsample(self, n=1):
        """``mu_lie @ rodrigues(v)`` for the cached algebra sample (``reparameterize.py:269-273``): like the reference a
        pure function of the module's cached state -- no fresh noise, nothing overwritten (``forward`` itself obtains z
        from the fused kernel)."""
        if self.return_means:
            return self.mu_lie.expand(n, *[-1] * len(self.mu_lie.shape))
        return self.mu_lie @ rodrigues(self.v)

    def kl(self):
        log_q_z_x = self.log_posterior()
        log_p_z = self.log_prior()
        kl = log_q_z_x - log_p_z
        return kl.mean(0)

    def log_posterior(self):
        return self._log_q

    def log_prior(self):
        prior = torch.tensor([LOG_PRIOR_SO3], dtype=torch.float64, device=self.z.device)
        return prior.expand_as(self.z[..., 0, 0])

    def deterministic(self):
        """Set to return means."""
        self.return_means = True
        self.reparameterize.deterministic()
