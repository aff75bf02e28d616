"""The regularisers: oracle restatement and the CUDA path against ``tests/golden/losses.npz``, which was
written by the reference's own ``EquivarianceLoss`` / ``EncoderContinuityLoss`` classes (``tests/golden/make_golden_losses.py``)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import so3_oracle as O


def T(a, **kw):
    return torch.tensor(np.asarray(a), dtype=torch.float64, **kw)


def test_oracle_equivariance_matches_reference_class():
    g = load_golden("losses")
    enc, enc2 = T(g["enc"]).requires_grad_(True), T(g["enc2"]).requires_grad_(True)
    diffs = O.equivariance_sqdist(T(g["theta"]), enc, enc2)
    np.testing.assert_allclose(diffs.detach().numpy(), g["diffs"], rtol=1e-12, atol=1e-13)
    loss = diffs.mean() * float(g["lamb"])
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-13)
    loss.backward()
    np.testing.assert_allclose(enc.grad.numpy(), g["g_enc"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(enc2.grad.numpy(), g["g_enc2"], rtol=1e-11, atol=1e-13)


def test_continuity_loss_matches_reference_class():
    """Host-side arithmetic: runs on CPU tensors too."""
    from lie_vae_b200.losses import EncoderContinuityLoss
    g = load_golden("losses")
    pairs = T(g["pairs"]).requires_grad_(True)
    loss = EncoderContinuityLoss(None, lamb=float(g["c_lamb"]))(pairs, 0)
    np.testing.assert_allclose(float(loss), float(g["c_loss"]), rtol=1e-13)
    loss.backward()
    np.testing.assert_allclose(pairs.grad.numpy(), g["g_pairs"], rtol=1e-12, atol=1e-14)
    assert float(EncoderContinuityLoss(None, lamb=lambda it: 2.0)(torch.arange(12.0).view(4, 3), 0)) == 2.0 * 27.0


def test_equivariance_loss_needs_cuda():
    """No CPU fallback for the kernel-backed part."""
    from lie_vae_b200.losses import equivariance_sqdist
    R = O.random_group_matrices(3)
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        equivariance_sqdist(torch.zeros(3), R, R)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gpu_equivariance_matches_reference_class(dtype):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from lie_vae_b200.losses import EquivarianceLoss, equivariance_sqdist
    g = load_golden("losses")
    tol = dict(rtol=1e-11, atol=1e-12) if dtype == torch.float64 else dict(rtol=1e-5, atol=1e-5)
    dev = lambda a: torch.tensor(np.asarray(a), dtype=dtype, device="cuda")   # noqa: E731
    enc, enc2 = dev(g["enc"]).requires_grad_(True), dev(g["enc2"]).requires_grad_(True)
    diffs = equivariance_sqdist(dev(g["theta"]), enc, enc2)
    np.testing.assert_allclose(diffs.detach().double().cpu().numpy(), g["diffs"], **tol)
    (diffs.mean() * float(g["lamb"])).backward()
    np.testing.assert_allclose(enc.grad.double().cpu().numpy(), g["g_enc"], **tol)
    np.testing.assert_allclose(enc2.grad.double().cpu().numpy(), g["g_enc2"], **tol)

    # the module, driven like the reference's: the stored angles in place of torch.rand, a stand-in encoder
    class Model:
        def __init__(self, out):
            self.out = out

        def encode(self, x):
            return ((self.out,),)

    class Log:
        def __init__(self):
            self.rows = []

        def add_scalar(self, *a):
            self.rows.append(a)

    from unittest import mock
    rtol = 1e-5 if dtype == torch.float32 else 1e-11
    log = Log()
    mod = EquivarianceLoss(Model(enc2.detach()), lamb=lambda it: float(g["lamb"]), log=log, report_freq=1)
    u = dev(g["theta"]) / (2 * np.pi)
    img = torch.randn(u.shape[0], 1, 4, 4, device="cuda", dtype=dtype)
    with mock.patch.object(torch, "rand", lambda *a, **k: u.clone()):
        loss = mod(img, enc.detach(), 0)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=rtol)
    assert [r[0] for r in log.rows] == ["equivariance", "equivariance_lamb"] and mod.diffs == []
    # num_samples truncation, constant weight
    mod5 = EquivarianceLoss(Model(enc2.detach()[:5]), num_samples=5, lamb=1.0)
    with mock.patch.object(torch, "rand", lambda *a, **k: u[:5].clone()):
        loss5 = mod5(img, enc.detach(), 3)
    np.testing.assert_allclose(float(loss5), float(np.mean(g["diffs"][:5])), rtol=rtol)
    assert len(mod5.diffs) == 1 and mod5.diffs[0].shape == (5,)
    assert EquivarianceLoss.rotate(img, u * 0).sub(img).abs().max().item() < 1e-5      # theta = 0: the identity warp


@pytest.mark.gpu
def test_gpu_equivariance_large_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from lie_vae_b200.losses import equivariance_sqdist
    torch.manual_seed(3)
    n = 100003
    th = torch.rand(n, dtype=torch.float64) * 2 * np.pi
    R = O.random_group_matrices(n, dtype=torch.float64)
    R2 = O.random_group_matrices(n, dtype=torch.float64) + 0.1 * torch.randn(n, 3, 3, dtype=torch.float64)
    w = torch.randn(n, dtype=torch.float64)
    a, b = R.clone().requires_grad_(True), R2.clone().requires_grad_(True)
    ref = O.equivariance_sqdist(th, a, b)
    (ref * w).sum().backward()
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 1e-5)):
        c, d = R.to("cuda", dtype).requires_grad_(True), R2.to("cuda", dtype).requires_grad_(True)
        out = equivariance_sqdist(th.to("cuda", dtype), c, d)
        (out * w.to("cuda", dtype)).sum().backward()
        np.testing.assert_allclose(out.detach().double().cpu().numpy(), ref.detach().numpy(), rtol=tol, atol=tol * 10)
        np.testing.assert_allclose(c.grad.double().cpu().numpy(), a.grad.numpy(), rtol=tol, atol=tol * 10)
        np.testing.assert_allclose(d.grad.double().cpu().numpy(), b.grad.numpy(), rtol=tol, atol=tol * 10)
    # empty batch
    e = equivariance_sqdist(torch.zeros(0, device="cuda"), torch.zeros(0, 3, 3, device="cuda"), torch.zeros(0, 3, 3, device="cuda"))
    assert e.shape == (0,)
