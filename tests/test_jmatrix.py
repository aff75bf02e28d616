"""The regenerated Pinchon-Hoggan J matrices against closed forms and structure."""
import math

import numpy as np

from lie_vae_b200.jmatrix import j_matrix_np, j_table, j_offsets, real_sph_harm


def test_closed_forms():
    np.testing.assert_allclose(j_matrix_np(0), [[1.0]], atol=1e-14)
    np.testing.assert_allclose(j_matrix_np(1), [[0, -1, 0], [-1, 0, 0], [0, 0, 1]], atol=1e-14)
    J2 = np.zeros((5, 5))
    J2[0, 3] = J2[3, 0] = -1
    J2[1, 1] = 1
    J2[2, 2] = -0.5
    J2[2, 4] = J2[4, 2] = -math.sqrt(3) / 2
    J2[4, 4] = 0.5
    np.testing.assert_allclose(j_matrix_np(2), J2, atol=1e-14)
    J3 = np.zeros((7, 7))
    for (i, j), v in {(0, 3): math.sqrt(10) / 4, (0, 5): -math.sqrt(6) / 4, (1, 1): 1.0, (2, 3): math.sqrt(6) / 4,
                      (2, 5): math.sqrt(10) / 4, (4, 4): -0.25, (4, 6): -math.sqrt(15) / 4, (6, 6): 0.25}.items():
        J3[i, j] = J3[j, i] = v
    np.testing.assert_allclose(j_matrix_np(3), J3, atol=1e-14)


def test_structure():
    nnz = []
    for l in range(13):
        J = j_matrix_np(l)
        d = 2 * l + 1
        assert J.shape == (d, d)
        np.testing.assert_allclose(J, J.T, atol=1e-14)
        np.testing.assert_allclose(J @ J, np.eye(d), atol=1e-12)
        nnz.append(int((J != 0).sum()))
    assert nnz[:9] == [1, 3, 7, 13, 21, 31, 43, 57, 71]
    assert j_offsets(3) == [0, 1, 10, 35, 84]
    t = j_table(8)
    assert t.dtype == np.float32 and t.size == 969


def test_definition_on_fresh_points():
    # Y(g p) = J Y(p) for g: (x,y,z) -> (x,-z,-y), at points that were not used in the fit
    rng = np.random.RandomState(3)
    p = rng.normal(size=(50, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    gp = np.stack([p[:, 0], -p[:, 2], -p[:, 1]], 1)
    for l in range(9):
        np.testing.assert_allclose(real_sph_harm(l, gp), real_sph_harm(l, p) @ j_matrix_np(l).T, atol=1e-12)
    # orthonormality of the harmonics themselves (quadrature on a fine lattice)
    y1 = real_sph_harm(1, np.eye(3))
    np.testing.assert_allclose(np.abs(y1), math.sqrt(3 / (4 * math.pi)) * np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]]), atol=1e-14)


def test_generated_header_is_current():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "lie_vae_b200", "csrc", "wigner_gen.cuh")
    before = open(path).read()
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_wigner.py"), "8"], check=True, capture_output=True)
    assert open(path).read() == before, "wigner_gen.cuh is stale: run tools/gen_wigner.py"


def test_body_frame_generator_identity():
    """The Wigner backward kernels take the angle gradients from the body-frame generators of D = X(a) J X(b) J X(c):
    G_z = X'(0), G_y = J G_z J, G_x = [G_z, G_y] and
        D^-1 dD/dc = G_z,  D^-1 dD/db = -sin(c) G_x + cos(c) G_y,  D^-1 dD/da = sin(b) cos(c) G_x + sin(b) sin(c) G_y + cos(b) G_z
    (csrc/wigner.cu, tools/gen_wigner.py).  Checked here against central differences of D in float64 for every degree the
    unrolled kernels cover, together with the so(3) commutation relations that make the identity hold."""
    import numpy as np
    from lie_vae_b200.jmatrix import j_matrix_np

    def X(l, phi):
        d = 2 * l + 1
        M = np.zeros((d, d))
        for i in range(d):
            M[i, i] = np.cos((l - i) * phi)
            if i != l:
                M[i, 2 * l - i] = np.sin((l - i) * phi)
        return M

    rng = np.random.default_rng(0)
    for l in range(1, 9):
        d = 2 * l + 1
        J = np.array(j_matrix_np(l))
        Gz = np.zeros((d, d))
        for i in range(d):
            if i != l:
                Gz[i, 2 * l - i] = l - i
        Gy = J @ Gz @ J
        Gx = Gz @ Gy - Gy @ Gz
        for G in (Gx, Gy, Gz):
            assert np.abs(G + G.T).max() < 1e-12                      # antisymmetric
        assert np.abs((Gy @ Gx - Gx @ Gy) - Gz).max() < 1e-10         # [G_y, G_x] = G_z
        assert np.abs((Gx @ Gz - Gz @ Gx) - Gy).max() < 1e-10         # [G_x, G_z] = G_y

        def D(a, b, c):
            return X(l, a) @ J @ X(l, b) @ J @ X(l, c)
        a, b, c = rng.uniform(-3, 3, 3)
        h = 1e-6
        D0 = D(a, b, c)
        Wa = D0.T @ (D(a + h, b, c) - D(a - h, b, c)) / (2 * h)
        Wb = D0.T @ (D(a, b + h, c) - D(a, b - h, c)) / (2 * h)
        Wc = D0.T @ (D(a, b, c + h) - D(a, b, c - h)) / (2 * h)
        tol = 1e-8 * (l + 1) ** 2
        assert np.abs(Wc - Gz).max() < tol
        assert np.abs(Wb - (-np.sin(c) * Gx + np.cos(c) * Gy)).max() < tol
        assert np.abs(Wa - (np.sin(b) * np.cos(c) * Gx + np.sin(b) * np.sin(c) * Gy + np.cos(b) * Gz)).max() < tol
