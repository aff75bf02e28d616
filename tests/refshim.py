"""Import the UNMODIFIED reference (``/root/reference``) in this container.

Test infrastructure only.  The reference needs three third-party packages that
are absent offline; stand-ins are placed in ``sys.modules``:

* ``hyperspherical_vae_pytorch.distributions`` (only used by the out-of-scope vMF path),
* ``lie_learn...pinchon_hoggan_dense.Jd`` backed by the J-free construction of ``oracle/wigner_direct.py`` (scipy's
  spherical harmonics; NOT the product's ``lie_vae_b200/jmatrix.py``, which ``tests/test_wigner_independent.py`` compares
  with it at 1e-12 -- the committed fixtures were generated when the stand-in still used the product table; the two agree to
  2e-15, far below the 1e-11 the fixtures are held to),
* ``lie_learn.groups.SO3.change_coordinates`` (image loading only).

``/root/reference`` does not exist on the GPU box: everything that uses this
module skips there.  Nothing is copied from the reference; it is imported.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lie_vae"))


class _JTable:
    def __getitem__(self, l):
        import numpy as np
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle.wigner_direct import j_matrix_direct
        return np.array(j_matrix_direct(l))


def _install_shims():
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    hv = mod("hyperspherical_vae_pytorch")
    hvd = mod("hyperspherical_vae_pytorch.distributions")
    hvd.VonMisesFisher = type("VonMisesFisher", (), {})
    hvd.HypersphericalUniform = type("HypersphericalUniform", (), {})
    hv.distributions = hvd

    names = ["lie_learn", "lie_learn.representations", "lie_learn.representations.SO3",
             "lie_learn.representations.SO3.pinchon_hoggan",
             "lie_learn.representations.SO3.pinchon_hoggan.pinchon_hoggan_dense",
             "lie_learn.groups", "lie_learn.groups.SO3"]
    for n in names:
        mod(n)
    sys.modules[names[4]].Jd = _JTable()

    def change_coordinates(*a, **k):
        raise NotImplementedError("lie_learn stand-in")
    sys.modules["lie_learn.groups.SO3"].change_coordinates = change_coordinates


def load_reference(float64_j=True):
    """Returns the reference's (lie_tools, reparameterize, decoders) modules.

    With ``float64_j`` the reference's ``j_matrix`` (hard-coded float32,
    ``lie_tools.py:14``) is wrapped so that the table follows the dtype of a
    module-level switch ``lie_tools._j_dtype`` -- a monkey-patch in this
    process, not an edit of the reference.
    """
    if not reference_available():
        raise RuntimeError("reference tree not present")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch
    import lie_vae.lie_tools as lt
    import lie_vae.reparameterize as rp
    import lie_vae.decoders as dc

    if float64_j and not hasattr(lt, "_j_dtype"):
        lt._j_dtype = torch.float32
        jd = sys.modules["lie_learn.representations.SO3.pinchon_hoggan.pinchon_hoggan_dense"].Jd

        def j_matrix(l, device=None):       # lie_tools.py:10-14 with the dtype made a switch; same Jd stand-in as above
            return torch.tensor(jd[l], dtype=lt._j_dtype, device=torch.device(device))
        lt.j_matrix = j_matrix
    return lt, rp, dc
