"""Wigner-D matrices and the Pinchon-Hoggan J matrices WITHOUT any J table: directly from the definition of the
representation on real spherical harmonics.  TEST INFRASTRUCTURE (see so3_oracle.py): the oracle's own source of J and the
J-free check of ``wigner_d_matrix`` that SURVEY.md section 8c / App. A describe.

The reference builds ``D^l = X(a) J X(b) J X(c)`` (``lie_tools.py:211-223``) with J from ``lie_learn``'s table
(``lie_tools.py:10-14``), which is absent here.  By App. A of the survey ``wigner_d_matrix((a,b,c), l)`` is the matrix of
``R = Rz(a) Ry(b) Rz(c)`` on the degree-l real spherical harmonics under ``Y(R p) = D Y(p)``, in lie_learn's basis
(real, orthonormal, m = -l..l):  m < 0: i/sqrt2 (Y_l^m - (-1)^m Y_l^-m),  m = 0: Y_l^0,  m > 0: 1/sqrt2 (Y_l^-m + (-1)^m Y_l^m)
with Condon-Shortley complex Y_l^m.  Here the complex harmonics come from ``scipy.special.sph_harm_y`` -- an implementation
that shares nothing with ``lie_vae_b200/jmatrix.py`` -- and a representation matrix is the exact solution of the
over-determined linear system  Y(R p_i) = D Y(p_i)  over random unit vectors p_i.
"""
from functools import lru_cache

import numpy as np
from scipy.special import sph_harm_y


def real_sh(l, pts):
    """(K, 3) unit vectors -> (K, 2l+1) real spherical harmonics, column m + l, in the basis stated above."""
    pts = np.asarray(pts, dtype=np.float64)
    theta = np.arccos(np.clip(pts[:, 2], -1.0, 1.0))          # polar
    phi = np.arctan2(pts[:, 1], pts[:, 0])                    # azimuth
    out = np.empty((pts.shape[0], 2 * l + 1))
    for m in range(-l, l + 1):
        if m == 0:
            out[:, l] = sph_harm_y(l, 0, theta, phi).real
        elif m > 0:
            v = (sph_harm_y(l, -m, theta, phi) + (-1) ** m * sph_harm_y(l, m, theta, phi)) / np.sqrt(2.0)
            out[:, l + m] = v.real
        else:
            v = 1j * (sph_harm_y(l, m, theta, phi) - (-1) ** m * sph_harm_y(l, -m, theta, phi)) / np.sqrt(2.0)
            out[:, l + m] = v.real
    return out


def _points(k, seed):
    p = np.random.RandomState(seed).normal(size=(k, 3))
    return p / np.linalg.norm(p, axis=1, keepdims=True)


def representation(l, A, seed=0):
    """Matrix D with Y(A p) = D Y(p) for an orthogonal 3x3 ``A`` (proper or improper), by exact least squares."""
    p = _points(max(96, 8 * (2 * l + 1)), seed + 17 * l)
    y_p, y_ap = real_sh(l, p), real_sh(l, p @ np.asarray(A, dtype=np.float64).T)
    dt, res, rank, _ = np.linalg.lstsq(y_p, y_ap, rcond=None)
    assert rank == 2 * l + 1
    return dt.T


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rot_y(b):
    c, s = np.cos(b), np.sin(b)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def wigner_d_direct(alpha, beta, gamma, l):
    """D^l of Rz(alpha) Ry(beta) Rz(gamma) from the definition -- no J anywhere."""
    return representation(l, rot_z(alpha) @ rot_y(beta) @ rot_z(gamma))


@lru_cache(maxsize=None)
def j_matrix_direct(l):
    """J_l = matrix of g: (x, y, z) -> (x, -z, -y) on the degree-l real harmonics (the oracle's own J source)."""
    g = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, -1.0, 0.0]])
    j = representation(int(l), g)
    j = 0.5 * (j + j.T)
    j[np.abs(j) < 1e-12] = 0.0
    j.setflags(write=False)
    return j
