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
