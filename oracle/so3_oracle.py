"""CPU oracle for the SO(3) latent hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain PyTorch (CPU, dtype-following: run it in float64 for parity checks)
restatement of the reference's algorithm for every function on the hot path
(SURVEY.md section 8a).  Each function cites the reference lines it follows.
Gradients come from autograd over these restatements, which is exactly how the
reference obtains them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Nothing under
``lie_vae_b200/`` imports it; the product path has no CPU fallback.

Pinning: ``tests/golden/*.npz`` were produced by importing the unmodified
reference from ``/root/reference`` (``tests/golden/make_golden.py``, float64,
values and autograd gradients); ``tests/test_oracle_golden.py`` checks this
module against them to 1e-12.  The Wigner J table is the one third-party input
the reference does not ship (lie_learn, unpinned, absent): with respect to
lie_learn's own table the Wigner values are *parity unpinned*; they are pinned
instead by the definition of the representation itself: ``oracle/wigner_direct.py``
builds D^l(a,b,c) for l <= 8 from scipy's spherical harmonics with no J anywhere
(``Y(R p) = D Y(p)``, the real basis lie_learn documents) and
``tests/test_wigner_independent.py`` holds ``wigner_d_matrix`` to it at 1e-10; the
J used here comes from the same J-free construction, not from the product's
``lie_vae_b200/jmatrix.py`` (the two are compared at 1e-12).  Also: closed-form
J_0..J_3, J=J^T, J^2=I and the reference's own orthogonality / inverse /
anti-homomorphism tests (``lie_tools.py:337-357``).
"""
import math

import numpy as np
import torch

TWO_PI = 2.0 * math.pi
LOG_PRIOR_SO3 = -math.log(8.0 * math.pi ** 2)      # reparameterize.py:266


# --------------------------------------------------------------------------- algebra maps
def map_to_lie_algebra(v):
    """hat map, lie_tools.py:17-43: [[0,-v2,v1],[v2,0,-v0],[-v1,v0,0]]."""
    assert v.shape[-1] == 3
    v0, v1, v2 = v[..., 0], v[..., 1], v[..., 2]
    o = torch.zeros_like(v0)
    rows = [torch.stack([o, -v2, v1], -1),
            torch.stack([v2, o, -v0], -1),
            torch.stack([-v1, v0, o], -1)]
    return torch.stack(rows, -2)


def map_to_lie_vector(X):
    """vee map, lie_tools.py:46-53."""
    return torch.stack([-X[..., 1, 2], X[..., 0, 2], -X[..., 0, 1]], -1)


def rodrigues(v):
    """exp map, lie_tools.py:56-64 (no small-angle guard: NaN at v = 0)."""
    theta = torch.linalg.vector_norm(v, dim=-1, keepdim=True)
    K = map_to_lie_algebra(v / theta)
    eye = torch.eye(3, dtype=v.dtype, device=v.device)
    th = theta[..., None]
    return eye + torch.sin(th) * K + (1.0 - torch.cos(th)) * (K @ K)


def s2s1rodrigues(s2_el, s1_el):
    """lie_tools.py:67-78: axis = s2_el (used as given), (cos, sin) = s1_el."""
    K = map_to_lie_algebra(s2_el)
    c = s1_el[..., 0, None, None]
    s = s1_el[..., 1, None, None]
    eye = torch.eye(3, dtype=s2_el.dtype, device=s2_el.device)
    return eye + s * K + (1.0 - c) * (K @ K)


def equivariance_sqdist(theta, encoding, encoding_of_rotated):
    """The SO(3) part of EquivarianceLoss.forward, losses/equivariance_loss.py:27-36: g = s2s1rodrigues(e_x, (cos, sin)),
    enc_rot = g.bmm(encoding), diffs = (enc_rot - img_rot_enc).pow(2).view(n, -1).sum(-1).  (n),(n,3,3),(n,3,3)->(n)."""
    n = theta.shape[0]
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=encoding.dtype, device=encoding.device).unsqueeze(0).expand(n, 3)
    g = s2s1rodrigues(ex, torch.stack((torch.cos(theta), torch.sin(theta)), 1))
    return (g.bmm(encoding) - encoding_of_rotated).pow(2).reshape(n, -1).sum(-1)


def s2s2_gram_schmidt(v1, v2):
    """lie_tools.py:81-89: rows e1, e2, e1 x e2; norms clamped at 1e-5.  (N,3),(N,3)->(N,3,3).

    The reference calls ``torch.cross`` without ``dim`` (first size-3 axis); for
    the (N,3) inputs it is used with and N != 3 that is the last axis, which is
    what is restated here.
    """
    e1 = v1 / torch.linalg.vector_norm(v1, dim=-1, keepdim=True).clamp(min=1e-5)
    u2 = v2 - (e1 * v2).sum(-1, keepdim=True) * e1
    e2 = u2 / torch.linalg.vector_norm(u2, dim=-1, keepdim=True).clamp(min=1e-5)
    e3 = torch.linalg.cross(e1, e2, dim=-1)
    return torch.stack([e1, e2, e3], 1)


def vector_to_eazyz(v):
    """lie_tools.py:92-97."""
    scale = v.new_tensor([math.pi, math.pi / 2, math.pi])
    shift = v.new_tensor([0.0, math.pi / 2, 0.0])
    return torch.tanh(v) * scale + shift


def log_map(R):
    """lie_tools.py:100-109, batched over leading dims (the reference is single-matrix)."""
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    theta = torch.acos(0.5 * (tr - 1.0))
    f = (theta / torch.sin(theta))[..., None, None]
    return f * (0.5 * (R - R.transpose(-1, -2)))


# --------------------------------------------------------------------------- coordinates
def group_matrix_to_quaternions(r):
    """lie_tools.py:112-157: four Shepperd candidates, pick argmax of the detached denominators."""
    lead = r.shape[:-2]
    assert tuple(r.shape[-2:]) == (3, 3)
    m = r.reshape(-1, 3, 3)
    d0, d1, d2 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    pre = torch.stack([1 + d0 - d1 - d2, 1 - d0 + d1 - d2, 1 - d0 - d1 + d2, 1 + d0 + d1 + d2], 1)
    den = 0.5 * torch.sqrt(1e-6 + pre.abs())
    s01, s02, s12 = m[:, 0, 1] + m[:, 1, 0], m[:, 0, 2] + m[:, 2, 0], m[:, 1, 2] + m[:, 2, 1]
    a12, a20, a01 = m[:, 1, 2] - m[:, 2, 1], m[:, 2, 0] - m[:, 0, 2], m[:, 0, 1] - m[:, 1, 0]
    f = 4 * den
    cand = torch.stack([
        torch.stack([den[:, 0], s01 / f[:, 0], s02 / f[:, 0], a12 / f[:, 0]], 1),
        torch.stack([s01 / f[:, 1], den[:, 1], s12 / f[:, 1], a20 / f[:, 1]], 1),
        torch.stack([s02 / f[:, 2], s12 / f[:, 2], den[:, 2], a01 / f[:, 2]], 1),
        torch.stack([a12 / f[:, 3], a20 / f[:, 3], a01 / f[:, 3], den[:, 3]], 1),
    ], 1)
    pick = den.detach().argmax(1)
    q = cand[torch.arange(m.shape[0]), pick]
    return q.reshape(*lead, 4)


def quaternions_to_eazyz(q):
    """lie_tools.py:160-175 (angles not reduced mod 2 pi; acos argument clamped to +-(1-1e-6))."""
    lead = q.shape[:-1]
    assert q.shape[-1] == 4
    q = q.reshape(-1, 4)
    q0, q1, q2, q3 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    eps = 1e-6
    alpha = torch.atan2(q1 * q2 - q0 * q3, q0 * q2 + q1 * q3)
    beta = torch.acos(torch.clamp(q3 ** 2 - q0 ** 2 - q1 ** 2 + q2 ** 2, -1.0 + eps, 1.0 - eps))
    gamma = torch.atan2(q0 * q3 + q1 * q2, q1 * q3 - q0 * q2)
    return torch.stack([alpha, beta, gamma], 1).reshape(*lead, 3)


def group_matrix_to_eazyz(r):
    """lie_tools.py:178-180."""
    return quaternions_to_eazyz(group_matrix_to_quaternions(r))


def quaternions_to_group_matrix(q):
    """lie_tools.py:183-192: normalise, then nine quadratic forms."""
    q = q / torch.linalg.vector_norm(q, dim=-1, keepdim=True)
    a, b, c, d = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    flat = torch.stack([
        a * a - b * b - c * c + d * d, 2 * (a * b + c * d), 2 * (a * c - b * d),
        2 * (a * b - c * d), -a * a + b * b - c * c + d * d, 2 * (b * c + a * d),
        2 * (a * c + b * d), 2 * (b * c - a * d), -a * a - b * b + c * c + d * d], -1)
    return flat.reshape(*q.shape[:-1], 3, 3)


def random_quaternions(n, dtype=torch.float32, device=None, generator=None):
    """lie_tools.py:256-263."""
    u1, u2, u3 = torch.rand(3, n, dtype=dtype, device=device, generator=generator)
    return torch.stack([torch.sqrt(1 - u1) * torch.sin(TWO_PI * u2),
                        torch.sqrt(1 - u1) * torch.cos(TWO_PI * u2),
                        torch.sqrt(u1) * torch.sin(TWO_PI * u3),
                        torch.sqrt(u1) * torch.cos(TWO_PI * u3)], 1)


def random_group_matrices(n, dtype=torch.float32, device=None, generator=None):
    """lie_tools.py:266-267."""
    return quaternions_to_group_matrix(random_quaternions(n, dtype, device, generator))


# --------------------------------------------------------------------------- Wigner
def _j_np(l):
    # the oracle's own J source: the representation matrix of (x,y,z) -> (x,-z,-y) on scipy's spherical harmonics
    # (oracle/wigner_direct.py) -- independent of the product's lie_vae_b200/jmatrix.py, which tests compare it with.
    from . import wigner_direct
    return wigner_direct.j_matrix_direct(int(l))


def j_matrix(l, dtype=torch.float64, device=None):
    """lie_tools.py:10-14, but following ``dtype`` (the reference hard-codes float32)."""
    return torch.as_tensor(np.array(_j_np(l)), dtype=dtype, device=device)


def z_rot_mat(angle, l):
    """lie_tools.py:195-208: diag cos(m phi), anti-diag sin(m phi), m = l..-l by row."""
    d = 2 * l + 1
    out = angle.new_zeros((angle.shape[0], d, d))
    idx = torch.arange(d, device=angle.device)
    freq = torch.arange(l, -l - 1, -1, dtype=angle.dtype, device=angle.device)[None]
    arg = freq * angle[:, None]
    out[:, idx, d - 1 - idx] = torch.sin(arg)
    out[:, idx, idx] = torch.cos(arg)
    return out


def wigner_d_matrix(angles, degree):
    """lie_tools.py:211-223: D^l = X(a) J X(b) J X(c)."""
    lead = angles.shape[:-1]
    assert angles.shape[-1] == 3
    a = angles.reshape(-1, 3)
    J = j_matrix(degree, a.dtype, a.device)
    D = z_rot_mat(a[:, 0], degree) @ J @ z_rot_mat(a[:, 1], degree) @ J @ z_rot_mat(a[:, 2], degree)
    d = 2 * degree + 1
    return D.reshape(*lead, d, d)


def block_wigner_matrix_multiply(angles, spectrum, max_degree, transpose=False):
    """lie_tools.py:226-253: per-degree D^l (or its transpose) times the degree-l rows."""
    pieces, start = [], 0
    for l in range(max_degree + 1):
        d = 2 * l + 1
        D = wigner_d_matrix(angles, l)
        if transpose:
            D = D.transpose(-2, -1)
        pieces.append(torch.bmm(D, spectrum[:, start:start + d, :]))
        start += d
    return torch.cat(pieces, 1)


def action_net_forward(angles, item_rep, degrees, transpose=False):
    """decoders.py:47-56 without mlp/deconv: expand item_rep over the batch, act, flatten."""
    n = angles.shape[0]
    M = (degrees + 1) ** 2
    spec = item_rep.expand(n, -1, -1)
    return block_wigner_matrix_multiply(angles, spec, degrees, transpose).reshape(n, M * item_rep.shape[1])


# --------------------------------------------------------------------------- reparameterize
def logsumexp(inputs, dim=None, keepdim=False):
    """utils.py:4-26."""
    if dim is None:
        inputs, dim = inputs.reshape(-1), 0
    s = inputs.max(dim=dim, keepdim=True)[0]
    out = s + (inputs - s).exp().sum(dim=dim, keepdim=True).log()
    return out if keepdim else out.squeeze(dim)


def n0_sample(sigma, eps):
    """reparameterize.py:137-141 with the noise made explicit: v = eps * sigma, eps (n,B,3), sigma (B,3)."""
    return eps * sigma


def so3_sample(mu, v):
    """reparameterize.py:269-273: z = mu @ exp(v), mu (B,3,3) broadcast over n."""
    return mu @ rodrigues(v)


def so3_log_posterior(v, sigma, k):
    """reparameterize.py:233-263 (+ N0reparameterize._log_posterior :131-132, utils.logsumexp).

    v (n,B,3), sigma (B,3) -> (n,B).
    """
    theta = torch.linalg.vector_norm(v, dim=-1, keepdim=True)            # (n,B,1)
    u = v / theta
    shifts = TWO_PI * torch.arange(-k, k + 1, dtype=v.dtype, device=v.device)
    theta_hat = theta[..., None, :] + shifts[:, None]                    # (n,B,2k+1,1)
    x = u[..., None, :] * theta_hat                                      # (n,B,2k+1,3)
    sg = sigma[None, :, None, :]
    log_n = (-(x ** 2) / (2 * sg ** 2) - sg.log() - 0.5 * math.log(TWO_PI)).sum(-1)   # (n,B,2k+1)
    clamp = 1e-3
    num = torch.clamp(theta_hat ** 2, min=clamp)
    den = torch.clamp(2 - 2 * torch.cos(theta_hat), min=clamp)
    log_vol = torch.log(num / den).sum(-1)
    return logsumexp(log_n + log_vol, -1)


def so3_log_prior(z):
    """reparameterize.py:265-267: constant -log(8 pi^2), float64, shape z[...,0,0]."""
    return torch.full(z.shape[:-2], LOG_PRIOR_SO3, dtype=torch.float64, device=z.device)


def so3_reparameterize(mu, sigma, eps, k):
    """The fused unit the CUDA kernel replaces: (z, log_q) from (mu, sigma, eps)."""
    v = n0_sample(sigma, eps)
    return so3_sample(mu, v), so3_log_posterior(v, sigma, k)


def so3_kl(log_q):
    """reparameterize.py:227-231."""
    return (log_q - LOG_PRIOR_SO3).mean(0)
