"""Wigner-D values pinned WITHOUT the J table: the reference's formula
``D^l = X(a) J X(b) J X(c)`` (``lie_tools.py:211-223``) against the representation matrix of ``Rz(a) Ry(b) Rz(c)`` on real
spherical harmonics built directly from ``scipy.special.sph_harm_y`` (oracle/wigner_direct.py).  A per-|m| sign or basis
error in the regenerated J for l >= 4 -- invisible to orthogonality / homomorphism tests -- fails here.  CPU only.
"""
import numpy as np
import pytest
import torch

from oracle import so3_oracle as O
from oracle import wigner_direct as W
from lie_vae_b200.jmatrix import j_matrix_np, real_sph_harm

ANGLES = [(0.3, 1.1, -2.0), (-2.7, 0.4, 1.9), (1.234, 2.9, 0.001), (0.0, 1.5707963267948966, 0.0), (3.0, 0.05, -3.0)]


@pytest.mark.parametrize("l", range(9))
def test_wigner_d_equals_direct_representation(l):
    for a, b, c in ANGLES:
        direct = W.wigner_d_direct(a, b, c, l)
        ours = O.wigner_d_matrix(torch.tensor([[a, b, c]], dtype=torch.float64), l)[0].numpy()
        np.testing.assert_allclose(ours, direct, rtol=0, atol=1e-10, err_msg="l=%d angles=%s" % (l, (a, b, c)))


def test_product_j_equals_independent_j():
    """lie_vae_b200/jmatrix.py (what the kernels bake in) against the J-free construction, and the two harmonic
    implementations against each other, beyond the degrees the fast kernels use."""
    p = W._points(40, 5)
    for l in range(13):
        np.testing.assert_allclose(j_matrix_np(l), W.j_matrix_direct(l), rtol=0, atol=1e-12)
        np.testing.assert_allclose(real_sph_harm(l, p), W.real_sh(l, p), rtol=0, atol=1e-12)


def test_product_j_in_the_chain_equals_direct_representation():
    """The same check with the PRODUCT's J in the chain (numpy, no oracle): X(a) J X(b) J X(c) == direct D."""
    def xmat(phi, l):
        d = 2 * l + 1
        x = np.zeros((d, d))
        idx = np.arange(d)
        freq = np.arange(l, -l - 1, -1)
        x[idx, d - 1 - idx] = np.sin(freq * phi)
        x[idx, idx] = np.cos(freq * phi)            # lie_tools.py:195-208 (the centre element ends up 1)
        return x
    for l in (4, 5, 6, 7, 8):
        J = j_matrix_np(l)
        for a, b, c in ANGLES[:3]:
            chain = xmat(a, l) @ J @ xmat(b, l) @ J @ xmat(c, l)
            np.testing.assert_allclose(chain, W.wigner_d_direct(a, b, c, l), rtol=0, atol=1e-10)


def test_l1_block_is_the_rotation():
    """D^1 in the (y, z, x) basis is the rotation matrix itself."""
    a, b, c = ANGLES[0]
    R = W.rot_z(a) @ W.rot_y(b) @ W.rot_z(c)
    P = np.array([[0.0, 1, 0], [0, 0, 1], [1, 0, 0]])
    np.testing.assert_allclose(W.wigner_d_direct(a, b, c, 1), P @ R @ P.T, atol=1e-12)
