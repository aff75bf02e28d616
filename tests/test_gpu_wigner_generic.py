"""GPU parity for the generic Wigner-D kernels (any degree <= 32, float32 / float64; csrc/wigner_generic.cu).

Through the public API (lie_tools / decoders, which dispatch on dtype and degree) and the C ABI underneath:
float64 against the reference's fixtures at 1e-9, float32 above degree 8 against the same fixtures at the
north-star tolerance, generic == unrolled kernels where both apply, and the reference's own group-property
tests (lie_tools.py:337-357) at degree 16.  ``-m gpu``.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_parity import RTOL, ATOL, close, dev

pytestmark = pytest.mark.gpu

F64TOL = dict(rtol=1e-9, atol=1e-10)


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import lie_vae_b200.lie_tools as lt
    import lie_vae_b200.decoders as dc
    from lie_vae_b200 import _ops
    return lt, dc, _ops


def dev64(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda")


def close64(x, ref, what=""):
    np.testing.assert_allclose(x.detach().cpu().numpy(), ref, err_msg=what, **F64TOL)


def test_wigner_d_f64_and_high_degree(mods):
    lt, _, _ = mods
    g = load_golden("wigner_d")
    for l in range(9):
        close64(lt.wigner_d_matrix(dev64(g["angles"]), l), g["D%d" % l], "D%d" % l)
    g = load_golden("wigner_d_high")
    for l in (9, 12, 16):
        close64(lt.wigner_d_matrix(dev64(g["angles"]), l), g["D%d" % l], "D%d f64" % l)
        close(lt.wigner_d_matrix(dev(g["angles"]), l), g["D%d" % l], RTOL, ATOL, "D%d f32" % l)


@pytest.mark.parametrize("tag", ["L8C3", "L3C1", "L5C10", "L11C2"])
@pytest.mark.parametrize("tr", ["N", "T"])
def test_block_wigner_f64_matches_reference(mods, tag, tr):
    lt, _, _ = mods
    g = load_golden("block_wigner_%s_%s" % (tag, tr))
    L = int(g["max_degree"])
    a, s = dev64(g["angles"]).requires_grad_(True), dev64(g["spectrum"]).requires_grad_(True)
    out = lt.block_wigner_matrix_multiply(a, s, L, transpose=(tr == "T"))
    assert out.dtype == torch.float64
    (out * dev64(g["w"])).sum().backward()
    close64(out, g["out"])
    close64(s.grad, g["gspectrum"])
    close64(a.grad, g["gangles"])


@pytest.mark.parametrize("tr", ["N", "T"])
def test_block_wigner_f32_degree_11(mods, tr):
    lt, _, _ = mods
    g = load_golden("block_wigner_L11C2_%s" % tr)
    a, s = dev(g["angles"]).requires_grad_(True), dev(g["spectrum"]).requires_grad_(True)
    out = lt.block_wigner_matrix_multiply(a, s, 11, transpose=(tr == "T"))
    (out * dev(g["w"])).sum().backward()
    close(out, g["out"], RTOL, ATOL)
    close(s.grad, g["gspectrum"], RTOL, ATOL)
    close(a.grad, g["gangles"], 2e-5, 5e-5)


@pytest.mark.parametrize("name,dtype", [("action_net_L10C4", torch.float32), ("action_net_L10C4", torch.float64),
                                        ("action_net_L8C10", torch.float64), ("action_net_L3C3", torch.float64)])
def test_action_net_generic(mods, name, dtype):
    _, dc, _ = mods
    g = load_golden(name)
    L, tr = int(g["degrees"]), bool(int(g["transpose"]))
    C = g["item_rep"].shape[1]
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, transpose=tr).cuda().to(dtype)
    net.item_rep.data = torch.tensor(g["item_rep"], dtype=dtype, device="cuda")
    a = torch.tensor(g["angles"], dtype=dtype, device="cuda").requires_grad_(True)
    out = net(a)
    (out * torch.tensor(g["w"], dtype=dtype, device="cuda")).sum().backward()
    if dtype == torch.float64:
        close64(out, g["out"]); close64(net.item_rep.grad, g["gitem"]); close64(a.grad, g["gangles"])
    else:
        scale = max(1.0, float(np.abs(g["gitem"]).max()))
        close(out, g["out"], RTOL, ATOL)
        close(net.item_rep.grad, g["gitem"], RTOL, ATOL * scale)
        close(a.grad, g["gangles"], 2e-5, 1e-4)


@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("tr", [False, True])
def test_generic_equals_unrolled_kernels(mods, shared, tr):
    """Where both paths apply (float32, degrees <= 8) they agree to rounding."""
    _, _, ops = mods
    torch.manual_seed(3)
    N, L, C = 777, 8, 10
    M = (L + 1) ** 2
    ang = (torch.rand(N, 3, device="cuda") * 6.0 - 3.0)
    spec = torch.randn(M, C, device="cuda") if shared else torch.randn(N, M, C, device="cuda")
    w = torch.randn(N, M, C, device="cuda")
    res = []
    for fn in (ops.WignerApply, ops.WignerApplyGeneric):
        a, s = ang.clone().requires_grad_(True), spec.clone().requires_grad_(True)
        out = fn.apply(a, s, 0, L, tr)
        (out * w).sum().backward()
        res.append((out.detach(), a.grad, s.grad))
    (o1, ga1, gs1), (o2, ga2, gs2) = res
    assert (o1 - o2).abs().max().item() <= 2e-5
    assert (ga1 - ga2).abs().max().item() <= 1e-3 * max(1.0, ga1.abs().max().item())
    assert (gs1 - gs2).abs().max().item() <= 2e-5 * max(1.0, gs1.abs().max().item())


def test_group_properties_degree_16(mods):
    """lie_tools.py:337-357 (orthogonality, inverse, anti-homomorphism) at a degree only the generic path has."""
    lt, _, _ = mods
    torch.manual_seed(4)
    # orthogonality holds for any angles (tight); inverse / product go through matrix -> Euler, whose 1e-6 guards
    # (lie_tools.py:120-126, 167-170) perturb the angles by ~1e-6/sin(beta): the reference tests them at 1e-3
    for dtype, tol, tol_g in ((torch.float64, 1e-10, 1e-3), (torch.float32, 2e-4, 1e-3)):
        r1 = lt.random_group_matrices(20, dtype=dtype, device="cuda")
        r2 = lt.random_group_matrices(20, dtype=dtype, device="cuda")
        a1, a2, a12 = (lt.group_matrix_to_eazyz(r) for r in (r1, r2, r1 @ r2))
        a1i = lt.group_matrix_to_eazyz(r1.transpose(-1, -2))
        for l in (9, 16):
            d1, d2, d12, d1i = (lt.wigner_d_matrix(a, l) for a in (a1, a2, a12, a1i))
            eye = torch.eye(2 * l + 1, dtype=dtype, device="cuda")
            assert (d1 @ d1.transpose(-1, -2) - eye).abs().max().item() < tol
            assert (d1 @ d1i - eye).abs().max().item() < tol_g
            assert (d2 @ d1 - d12).abs().max().item() < tol_g        # W(b) W(a) = W(ab)


def test_generic_errors(mods):
    lt, dc, ops = mods
    with pytest.raises(NotImplementedError):
        lt.wigner_d_matrix(torch.zeros(2, 3, device="cuda"), 33)
    with pytest.raises(NotImplementedError):
        dc.ActionNet(33, torch.nn.Sequential())
    with pytest.raises(TypeError):
        ops.WignerApplyGeneric.apply(torch.zeros(2, 3, device="cuda"), torch.zeros(4, 2, device="cuda", dtype=torch.float64), 0, 1, False)
    out = lt.block_wigner_matrix_multiply(torch.empty(0, 3, device="cuda", dtype=torch.float64),
                                          torch.empty(0, 100, 2, device="cuda", dtype=torch.float64), 9)
    assert out.shape == (0, 100, 2)
