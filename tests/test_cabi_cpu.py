"""CPU-side checks of the boundary: the library loads, exports every symbol include/lievae.h declares,
argument errors are reported without touching a GPU, the host modules keep the reference's surface, and the
product path refuses CPU tensors (no fallback).  No compute calls here."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def lib():
    from lie_vae_b200 import _build, _cabi
    _build.build()          # no-op when the in-tree .so is current
    return _cabi.lib()


def test_header_symbols_exported(lib):
    from lie_vae_b200 import _cabi
    protos = _cabi.header_prototypes()
    text = open(_cabi.HEADER_PATH).read()
    declared = set(re.findall(r"\b(lv_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(protos), "header parser and header disagree"
    assert len(protos) >= 54
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.lv_version() == 100
    assert isinstance(_cabi.last_error(), str)


def test_argument_errors_without_gpu(lib):
    from lie_vae_b200 import _cabi
    # negative sizes / null pointers are rejected before any CUDA call
    assert lib.lv_rodrigues_fwd_f32(None, None, -1, None) == -1
    assert "negative" in _cabi.last_error()
    assert lib.lv_rodrigues_fwd_f32(None, None, 5, None) == -1
    assert "null" in _cabi.last_error()
    assert lib.lv_rodrigues_fwd_f64(None, None, 0, None) == 0
    assert lib.lv_so3_reparam_fwd_f32(None, None, None, None, None, 1, 4, 100, None) == -2
    assert lib.lv_so3_reparam_fwd_f32(None, None, None, None, None, 1, 4, 3, None) == -1
    assert lib.lv_so3_reparam_bwd_f32(None, None, None, None, None, None, None, 0, 4, 3, None) == 0
    assert lib.lv_sum_leading_f32(None, None, -1, 3, None) == -1
    with pytest.raises(RuntimeError, match="argument error"):
        _cabi.call("lv_hat_fwd_f32", None, None, 3, None)


def test_no_cpu_fallback():
    import lie_vae_b200.lie_tools as lt
    import lie_vae_b200.reparameterize as rp
    import lie_vae_b200.decoders as dc
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        lt.rodrigues(torch.randn(4, 3))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        rp.so3_reparameterize(torch.eye(3).expand(2, 3, 3), torch.ones(2, 3), torch.zeros(1, 2, 3), 3)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        dc.ActionNet(2, torch.nn.Sequential())(torch.zeros(4, 3))
    # the product package never imports the oracle
    for mod in ("lie_tools", "reparameterize", "decoders", "_ops", "_cabi", "pipeline", "dist", "utils", "jmatrix"):
        src = open(os.path.join(ROOT, "lie_vae_b200", mod + ".py")).read()
        assert "oracle" not in src.replace("no CPU", ""), mod


def test_module_surface_matches_reference():
    import lie_vae_b200.lie_tools as lt
    import lie_vae_b200.reparameterize as rp
    import lie_vae_b200.decoders as dc
    for name in ["map_to_lie_algebra", "map_to_lie_vector", "rodrigues", "s2s1rodrigues", "s2s2_gram_schmidt",
                 "vector_to_eazyz", "log_map", "group_matrix_to_quaternions", "quaternions_to_eazyz",
                 "group_matrix_to_eazyz", "quaternions_to_group_matrix", "wigner_d_matrix",
                 "block_wigner_matrix_multiply", "random_quaternions", "random_group_matrices", "j_matrix"]:
        assert callable(getattr(lt, name)), name
    m = rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.AlgebraMean(10), k=3)
    for attr in ["mean_module", "reparameterize", "input_dim", "k", "return_means", "mu_lie", "v", "z"]:
        assert hasattr(m, attr), attr
    assert m.input_dim == 10 and m.k == 3
    with pytest.raises(AssertionError):
        rp.SO3reparameterize(rp.N0reparameterize(10, 4), rp.AlgebraMean(10))
    assert sorted(rp.S2S1Mean(7).state_dict()) == ["s1_map.bias", "s1_map.weight", "s2_map.bias", "s2_map.weight"]
    assert "fixed_sigma" in rp.N0reparameterize(4, 3, fixed_sigma=0.5).state_dict()
    a = dc.ActionNet(3, torch.nn.Sequential(), rep_copies=4, with_mlp=True)
    for attr in ["degrees", "rep_copies", "matrix_dims", "transpose", "item_rep", "mlp", "deconv"]:
        assert hasattr(a, attr), attr
    assert a.matrix_dims == 16 and tuple(a.item_rep.shape) == (16, 4)
    assert [k for k in a.state_dict() if k.startswith("mlp.")] == ["mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias",
                                                                  "mlp.4.weight", "mlp.4.bias", "mlp.6.weight", "mlp.6.bias"]
    buf = dc.ActionNet(2, torch.nn.Sequential(), item_rep=torch.zeros(9, 10))
    assert "item_rep" in dict(buf.named_buffers()) and not list(buf.parameters())
    with pytest.raises(NotImplementedError):
        dc.ActionNet(33, torch.nn.Sequential())
    assert tuple(lt.j_matrix(3).shape) == (7, 7) and lt.j_matrix(3).dtype == torch.float32
    # every name experiments/vae.py:5-12 imports from reparameterize / decoders / lie_tools / utils resolves here
    for mod, names in ((rp, ["SO3reparameterize", "N0reparameterize", "Nreparameterize", "Sreparameterize", "AlgebraMean",
                             "QuaternionMean", "S2S1Mean", "S2S2Mean"]),
                       (dc, ["MLPNet", "ActionNet"]),
                       (lt, ["group_matrix_to_eazyz", "vector_to_eazyz", "quaternions_to_eazyz"])):
        for name in names:
            assert hasattr(mod, name), name
    import lie_vae_b200.utils as ut
    assert callable(ut.logsumexp)
    # the Euclidean baseline is host-side PyTorch: same state_dict keys and formulas as reparameterize.py:16-55
    torch.manual_seed(0)
    nr = rp.Nreparameterize(5, 3)
    assert sorted(nr.state_dict()) == ["mu_linear.bias", "mu_linear.weight", "sigma_linear.bias", "sigma_linear.weight"]
    z = nr(torch.randn(7, 5), n=4)
    assert tuple(z.shape) == (4, 7, 3) and tuple(nr.kl().shape) == (7,)
    ref = torch.distributions.Normal(nr.mu, nr.sigma).log_prob(z).sum(-1)
    assert torch.allclose(nr.log_posterior(), ref, atol=1e-6)
    ref0 = torch.distributions.Normal(torch.zeros_like(nr.mu), torch.ones_like(nr.sigma)).log_prob(z).sum(-1)
    assert torch.allclose(nr.log_prior(), ref0, atol=1e-6)
    nr.deterministic()
    assert torch.equal(nr.nsample(2)[1], nr.mu)
    sr = rp.Sreparameterize(5, 3)
    assert sorted(sr.state_dict()) == ["k_linear.bias", "k_linear.weight", "mu_linear.bias", "mu_linear.weight"]
    with pytest.raises(ImportError):
        sr(torch.randn(2, 5))
    mn = dc.MLPNet(2, torch.nn.Sequential(), in_dims=9, rep_copies=3)
    assert tuple(mn(torch.randn(4, 3, 3)).shape) == (4, 27)
    assert [k for k in mn.state_dict()][:2] == ["mlp.0.weight", "mlp.0.bias"]
