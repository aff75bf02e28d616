"""SO(3) group / algebra tools -- drop-in for the reference's ``lie_vae.lie_tools``.

Same names, argument meaning, shapes and error behaviour as ``lie_tools.py:10-267``
of the reference; every function is one hand-written sm_100a kernel (forward and
backward) behind the C ABI of ``include/lievae.h``.  CUDA tensors only.
"""
from functools import lru_cache
import math

import numpy as np
import torch

from . import _ops
from .jmatrix import j_matrix_np

__all__ = ["j_matrix", "map_to_lie_algebra", "map_to_lie_vector", "rodrigues", "s2s1rodrigues",
           "s2s2_gram_schmidt", "vector_to_eazyz", "log_map", "group_matrix_to_quaternions",
           "quaternions_to_eazyz", "group_matrix_to_eazyz", "quaternions_to_group_matrix",
           "wigner_d_matrix", "block_wigner_matrix_multiply", "random_quaternions", "random_group_matrices"]

MAX_DEGREE = 32   # degrees handled by the Wigner kernels (<= 8: unrolled packed float32 kernels; above, or float64: generic kernels)


@lru_cache(maxsize=256)
def j_matrix(l, device=None):
    """Pinchon-Hoggan J_l as a float32 tensor (``lie_tools.py:10-14``), regenerated analytically."""
    return torch.tensor(np.array(j_matrix_np(l)), dtype=torch.float32, device=torch.device(device or "cpu"))


def map_to_lie_algebra(v):
    """hat map R^3 -> so(3), (...,3) -> (...,3,3)   (``lie_tools.py:17-43``)."""
    assert v.size()[-1] == 3
    return _ops.Hat.apply(v)


def map_to_lie_vector(X):
    """vee map so(3) -> R^3, (...,3,3) -> (...,3)   (``lie_tools.py:46-53``)."""
    return _ops.Vee.apply(X)


def rodrigues(v):
    """Exponential map (...,3) -> (...,3,3)   (``lie_tools.py:56-64``).  v = 0 gives I (reference: NaN)."""
    return _ops.Rodrigues.apply(v)


def s2s1rodrigues(s2_el, s1_el):
    """Axis (...,3) and (cos, sin) (...,2) -> rotation (...,3,3)   (``lie_tools.py:67-78``)."""
    return _ops.S2S1Rodrigues.apply(s2_el, s1_el)


def s2s2_gram_schmidt(v1, v2):
    """Two 3-vectors (N,3) -> rotation (N,3,3) with rows e1, e2, e1 x e2   (``lie_tools.py:81-89``)."""
    return _ops.S2S2GramSchmidt.apply(v1, v2)


def vector_to_eazyz(v):
    """tanh squashing of a 3-vector to ZYZ Euler ranges   (``lie_tools.py:92-97``)."""
    return _ops.VectorToEazyz.apply(v)


def log_map(R):
    """Logarithm map (...,3,3) -> (...,3,3) algebra element   (``lie_tools.py:100-109``, batched)."""
    return _ops.LogMap.apply(R)


def group_matrix_to_quaternions(r):
    """(...,3,3) -> (...,4), scalar-last, Shepperd with argmax branch   (``lie_tools.py:112-157``)."""
    assert list(r.shape[-2:]) == [3, 3], 'Input must be 3x3 matrices'
    return _ops.MatToQuat.apply(r)


def quaternions_to_eazyz(q):
    """(...,4) -> (...,3) ZYZ Euler angles, not reduced mod 2 pi   (``lie_tools.py:160-175``)."""
    assert q.shape[-1] == 4, 'Input must be 4 dim vectors'
    return _ops.QuatToEazyz.apply(q)


def group_matrix_to_eazyz(r):
    """(...,3,3) -> (...,3), one fused kernel   (``lie_tools.py:178-180``)."""
    assert list(r.shape[-2:]) == [3, 3], 'Input must be 3x3 matrices'
    return _ops.MatToEazyz.apply(r)


def quaternions_to_group_matrix(q):
    """Normalises q and maps to a rotation matrix, (...,4) -> (...,3,3)   (``lie_tools.py:183-192``)."""
    return _ops.QuatToMat.apply(q)


def _z_rot_mat(angle, l):
    """X_l(angle): diag cos(m*angle), anti-diag sin(m*angle), m = l..-l by row   (``lie_tools.py:195-208``).

    Private helper of the reference (only ``wigner_d_matrix`` uses it); provided for completeness as
    D^l(angle, 0, 0) = X(angle) J X(0) J X(0) = X(angle).
    """
    zeros = torch.zeros_like(angle)
    return wigner_d_matrix(torch.stack([angle, zeros, zeros], -1), l)


def _check_degree(l):
    if l > MAX_DEGREE:
        raise NotImplementedError("degree %d > %d is not supported by the sm_100a Wigner kernels" % (l, MAX_DEGREE))


def wigner_d_matrix(angles, degree):
    """Wigner D matrices (...,3) -> (...,2l+1,2l+1) for ZYZ Euler angles   (``lie_tools.py:211-223``).

    Evaluated as the degree-l action on the identity spectrum (differentiable in the angles).
    """
    batch_dims = angles.shape[:-1]
    assert angles.shape[-1] == 3, 'Input must be 3 dim vectors'
    _check_degree(degree)
    d = 2 * degree + 1
    eye = torch.eye(d, dtype=angles.dtype, device=angles.device)
    out = _ops.wigner_apply(angles.reshape(-1, 3), eye, degree, degree, False)
    return out.view(*batch_dims, d, d)


def block_wigner_matrix_multiply(angles, spectrum, max_degree, transpose=False):
    """Act with D^0 + ... + D^L on a spectrum   (``lie_tools.py:226-253``).

    angles (batch,3); spectrum (batch,(L+1)^2,channels); returns the same shape.  A spectrum that is
    a stride-0 expand of one (M,C) matrix (``decoders.py:53``, ``datasets.py:153``) is detected and
    never materialised.
    """
    _check_degree(max_degree)
    if spectrum.dim() == 3 and spectrum.shape[0] != 1 and spectrum.stride(0) == 0:
        spectrum = spectrum[0]
    return _ops.wigner_apply(angles, spectrum, 0, max_degree, transpose)


def random_quaternions(n, dtype=torch.float32, device=None):
    """Uniform unit quaternions (n,4)   (``lie_tools.py:256-263``)."""
    u1, u2, u3 = torch.rand(3, n, dtype=dtype, device=device)
    return torch.stack((
        torch.sqrt(1 - u1) * torch.sin(2 * math.pi * u2),
        torch.sqrt(1 - u1) * torch.cos(2 * math.pi * u2),
        torch.sqrt(u1) * torch.sin(2 * math.pi * u3),
        torch.sqrt(u1) * torch.cos(2 * math.pi * u3),
    ), 1)


def random_group_matrices(n, dtype=torch.float32, device=None):
    """Uniform rotation matrices (n,3,3)   (``lie_tools.py:266-267``)."""
    return quaternions_to_group_matrix(random_quaternions(n, dtype, device))
