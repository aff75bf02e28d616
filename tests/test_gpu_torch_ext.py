"""The C++ autograd binding (csrc/torch_binding.cpp, the float32 CUDA fast path of _ops.py) against the Python Functions over
the same C ABI: bit-identical outputs and gradients, same argument errors.  ``-m gpu``."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from lie_vae_b200 import _ops
    if _ops.torch_ext() is None:
        pytest.fail("the C++ autograd extension is not built (python -m lie_vae_b200._build): the drop-in API would run on the "
                    "slower Python Functions")
    return _ops


@pytest.mark.parametrize("n,B,k", [(1, 1000, 3), (3, 257, 10), (1, 1 << 18, 3)])
@pytest.mark.parametrize("euler", [False, True])
def test_reparam_ext_equals_python_function(ops, n, B, k, euler):
    import lie_vae_b200.lie_tools as lt
    torch.manual_seed(B)
    mu = lt.random_group_matrices(B, device="cuda")
    sg = torch.nn.functional.softplus(torch.randn(B, 3, device="cuda"))
    eps = torch.randn(n, B, 3, device="cuda")
    wp = torch.randn((n, B, 3) if euler else (n, B, 3, 3), device="cuda")
    wl = torch.randn(n, B, device="cuda")
    res = []
    for path in ("ext", "python"):
        m, s = mu.clone().requires_grad_(True), sg.clone().requires_grad_(True)
        if path == "ext":
            pose, lq = ops.so3_reparam(m, s, eps, k, euler)
            assert "SO3ReparamFn" in pose.grad_fn.name()
        else:
            pose, lq = (ops.SO3ReparamEazyz if euler else ops.SO3Reparam).apply(m, s, eps, k)
        ((pose * wp).sum() + (lq * wl).sum()).backward()
        res.append((pose.detach(), lq.detach(), m.grad, s.grad))
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # only one output used: the other gradient arrives undefined, not as zeros
    m = mu.clone().requires_grad_(True)
    pose, lq = ops.so3_reparam(m, sg, eps, k, euler)
    lq.sum().backward()
    assert torch.isfinite(m.grad).all()


@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("L,C,N,tr", [(8, 10, 4099, False), (6, 10, 1000, True), (3, 2, 77, False)])
def test_wigner_ext_equals_python_function(ops, shared, L, C, N, tr):
    torch.manual_seed(N)
    M = (L + 1) ** 2
    ang = torch.rand(N, 3, device="cuda") * 6 - 3
    spec = torch.randn(M, C, device="cuda") if shared else torch.randn(N, M, C, device="cuda")
    g = torch.randn(N, M, C, device="cuda")
    res = []
    for path in ("ext", "python"):
        a, s = ang.clone().requires_grad_(True), spec.clone().requires_grad_(True)
        out = ops.wigner_apply(a, s, 0, L, tr) if path == "ext" else ops.WignerApply.apply(a, s, 0, L, tr)
        if path == "ext":
            assert "WignerApplyFn" in out.grad_fn.name()
        out.backward(g)
        res.append((out.detach(), a.grad, s.grad))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_ext_argument_errors(ops):
    a = torch.randn(4, 3, device="cuda")
    ext = ops.torch_ext()
    # the extension's own checks (a C++ exception must surface as a Python error, not as a crash) ...
    with pytest.raises(RuntimeError, match="spectrum must have 9 rows"):
        ext.wigner_apply(a, torch.randn(10, 2, device="cuda"), 0, 2, False)
    with pytest.raises(RuntimeError, match="sigma must be"):
        ext.so3_reparam(torch.randn(5, 3, 3, device="cuda"), torch.randn(4, 3, device="cuda"), torch.randn(1, 5, 3, device="cuda"), 3, False)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ext.so3_reparam(torch.randn(5, 3, 3), torch.randn(5, 3), torch.randn(1, 5, 3), 3, False)
    # ... and the public wrappers keep the Python Functions' error types
    with pytest.raises(ValueError):
        ops.wigner_apply(a, torch.randn(10, 2, device="cuda"), 0, 2, False)
    with pytest.raises(ValueError):
        ops.so3_reparam(torch.randn(5, 3, 3, device="cuda"), torch.randn(4, 3, device="cuda"), torch.randn(1, 5, 3, device="cuda"), 3)
    with pytest.raises(RuntimeError):                                                # CPU tensors: no fallback on either path
        ops.so3_reparam(torch.randn(5, 3, 3), torch.randn(5, 3), torch.randn(1, 5, 3), 3)
