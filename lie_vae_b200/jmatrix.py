"""Pinchon-Hoggan J matrices, regenerated analytically (host side, float64).

The reference obtains ``J_l`` from a third-party table
(``lie_tools.py:10-14`` -> ``lie_learn ... pinchon_hoggan_dense.Jd``), which is
not vendored and not installable offline.  ``J_l`` is fully determined by its
definition, so it is rebuilt here from first principles:

    J_l is the matrix, in the degree-l *real* spherical-harmonic basis
    (orthonormal, order m = -l..l, no Condon-Shortley sign in the real basis:
    Y_{1,-1} ~ y, Y_{1,0} ~ z, Y_{1,1} ~ x), of the point map
    g : (x, y, z) -> (x, -z, -y), acting as  Y(g p) = J_l Y(p).

``g`` swaps the z and y axes, which is what turns a z-rotation block into a
y-rotation block:  D(alpha, beta, gamma) = X(alpha) J X(beta) J X(gamma)
(``lie_tools.py:221``).  ``J_l`` is symmetric and an involution.

The matrices are obtained by evaluating the real harmonics on a fixed,
over-determined set of unit vectors and solving the (exactly consistent) linear
system in float64, then zeroing entries below 1e-12.  Known closed forms
(J_0..J_3) are checked in ``tests/test_jmatrix.py``.

Nothing here touches the GPU: the generic kernels take the packed table as a
caller-owned device pointer (``lv_wigner_generic_*``, ``_ops._j_table``), and
``tools/gen_wigner.py`` bakes it into ``csrc/wigner_gen.cuh`` for the unrolled,
packed ones.  Independent check: ``tests/test_wigner_independent.py``.
"""
from functools import lru_cache
import math

import numpy as np

__all__ = ["real_sph_harm", "j_matrix_np", "j_table", "j_offsets", "J_ZERO_TOL"]

J_ZERO_TOL = 1e-12


def real_sph_harm(l, pts):
    """Orthonormal real spherical harmonics of degree ``l``.

    pts: (K, 3) float64 unit vectors.  Returns (K, 2l+1), column m+l.
    m > 0: sqrt2 N P_l^m(z) cos(m phi); m < 0: sqrt2 N P_l^|m|(z) sin(|m| phi);
    P without the Condon-Shortley factor.
    """
    pts = np.asarray(pts, dtype=np.float64)
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    phi = np.arctan2(y, x)
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    # P[m][ll] for ll = m..l, recurrence in degree
    out = np.empty((pts.shape[0], 2 * l + 1), dtype=np.float64)
    pmm = np.ones_like(z)
    for m in range(0, l + 1):
        if m > 0:
            pmm = pmm * (2 * m - 1) * s
        # climb degree from m to l
        p_prev2 = pmm
        if l == m:
            plm = pmm
        else:
            p_prev1 = z * (2 * m + 1) * pmm
            for ll in range(m + 2, l + 1):
                p_cur = (z * (2 * ll - 1) * p_prev1 - (ll + m - 1) * p_prev2) / (ll - m)
                p_prev2, p_prev1 = p_prev1, p_cur
            plm = p_prev1
        norm = math.sqrt((2 * l + 1) / (4 * math.pi)
                         * math.factorial(l - m) / math.factorial(l + m))
        if m == 0:
            out[:, l] = norm * plm
        else:
            out[:, l + m] = math.sqrt(2.0) * norm * plm * np.cos(m * phi)
            out[:, l - m] = math.sqrt(2.0) * norm * plm * np.sin(m * phi)
    return out


def _sample_points(k):
    """Deterministic, well-spread unit vectors (Fibonacci lattice, offset to avoid poles)."""
    i = np.arange(k, dtype=np.float64) + 0.5
    z = 1.0 - 2.0 * i / k
    phi = i * math.pi * (3.0 - math.sqrt(5.0))
    r = np.sqrt(1.0 - z * z)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], 1)


@lru_cache(maxsize=None)
def _j_cached(l):
    k = max(64, 6 * (2 * l + 1))
    p = _sample_points(k)
    gp = np.stack([p[:, 0], -p[:, 2], -p[:, 1]], 1)
    y_p = real_sph_harm(l, p)
    y_gp = real_sph_harm(l, gp)
    # rows: Y(gp)^T = Y(p)^T J^T
    jt, *_ = np.linalg.lstsq(y_p, y_gp, rcond=None)
    j = jt.T
    j = 0.5 * (j + j.T)           # symmetric by construction; remove 1e-16 asymmetry
    j[np.abs(j) < J_ZERO_TOL] = 0.0
    j.setflags(write=False)
    return j


def j_matrix_np(l):
    """float64 ``J_l`` ((2l+1, 2l+1), read-only)."""
    if l < 0:
        raise ValueError("degree must be >= 0")
    return _j_cached(int(l))


def j_offsets(max_degree):
    """Start offset of each dense J_l block in the packed table (sum of (2l+1)^2)."""
    offs = [0]
    for l in range(max_degree + 1):
        offs.append(offs[-1] + (2 * l + 1) ** 2)
    return offs


def j_table(max_degree, dtype=np.float32):
    """Packed row-major dense blocks J_0 | J_1 | ... | J_L."""
    return np.concatenate([j_matrix_np(l).reshape(-1) for l in range(max_degree + 1)]).astype(dtype)
