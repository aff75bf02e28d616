"""The tensor-core consumer of the Wigner action: lv_gemm_tf32_f32 (tcgen05.mma kind::tf32, TMEM
accumulators, TMA weight tiles) and the fused ``ActionNet(fuse_consumer=True)`` path against the float64 oracle followed by
float64 ``conv_transpose2d`` / ``linear``.  Tolerance: TF32 operands have a 10-bit mantissa (2^-11 relative rounding per
operand, both operands rounded to nearest); with FP32 accumulation the error of an output is ~N(0, (0.6 * 2^-11 * rms)^2) whatever
K is, so the worst of up to 10^7 outputs stays below 3e-3 of the output's rms -- the arithmetic cuDNN uses for the reference's
FP32 ConvTranspose2d by default.  ``-m gpu``.
"""
import pytest
import torch

from oracle import so3_oracle as O

pytestmark = pytest.mark.gpu
TF32_TOL = 3e-3


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from lie_vae_b200 import _ops
    import lie_vae_b200.decoders as dc
    return _ops, dc


def rel_rms(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().pow(2).mean().sqrt().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 810), (1024, 3200, 810), (1000, 800, 810), (77, 50, 810), (300, 130, 64),
                                   (8192, 256, 490), (129, 3200, 810)])
def test_gemm_tf32_vs_float64(mods, M, N, K):
    _ops, _ = mods
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda")
    ldb = (K + 3) // 4 * 4
    bt = torch.zeros(N, ldb, device="cuda")
    bt[:, :K] = torch.randn(N, K, device="cuda")
    bias = torch.randn((N + 15) // 16, device="cuda")
    want = a.double() @ bt[:, :K].double().t() + bias.double().repeat_interleave(16)[:N]
    bt = _ops.round_tf32(bt)
    got = _ops.gemm_tf32(a, bt[:, :K], bias, 16)
    assert got.shape == (M, N) and torch.isfinite(got).all()
    assert rel_rms(got, want) < TF32_TOL
    # row-strided A (a view into a wider buffer) and no bias
    wide = torch.randn(M, K + 6, device="cuda")
    got2 = _ops.gemm_tf32(wide[:, :K], bt[:, :K])
    assert rel_rms(got2, wide[:, :K].double() @ bt[:, :K].double().t()) < TF32_TOL
    # exactness of the data path: with operands that TF32 represents exactly the result is the FP32 one up to summation order
    ai = torch.randint(-8, 9, (M, K), device="cuda").float()
    bi = torch.zeros(N, ldb, device="cuda")
    bi[:, :K] = torch.randint(-8, 9, (N, K), device="cuda").float()
    assert torch.equal(_ops.gemm_tf32(ai, bi[:, :K]), (ai.double() @ bi[:, :K].double().t()).float())


@pytest.mark.parametrize("wgrad_tf32", [True, False])
@pytest.mark.parametrize("L,Nout,div,N,tr,chunk", [(8, 800, 16, 1000, False, 8192), (6, 3200, 16, 1024, False, 300), (4, 50, 1, 777, True, 8192),
                                                   (8, 384, 16, 20001, False, 8192)])
def test_action_gemm_function_vs_float64(mods, L, Nout, div, N, tr, chunk, wgrad_tf32, monkeypatch):
    """ActionGemm (chunked Wigner forward -> tcgen05 GEMM; backward: recompute, tcgen05 dgrad, cuBLAS wgrad, Wigner backward)
    against float64 autograd of oracle-action @ weight + bias.  The output and the gradients that flow through the data-gradient
    GEMM (g_angles, g_item_rep) are TF32-accurate; the weight gradient is TF32-accurate by default (ACTION_GEMM_WGRAD_TF32, cuDNN's
    default for the reference's layer) and FP32-accurate with the switch off; the bias gradient is FP32."""
    _ops, _ = mods
    monkeypatch.setattr(_ops, "ACTION_GEMM_WGRAD_TF32", wgrad_tf32)
    allow_before = torch.backends.cuda.matmul.allow_tf32
    torch.manual_seed(L + Nout)
    C = 10
    Mh = (L + 1) ** 2
    K = Mh * C
    ang64 = O.group_matrix_to_eazyz(O.random_group_matrices(N, dtype=torch.float64)).requires_grad_(True)
    it64 = torch.randn(Mh, C, dtype=torch.float64, requires_grad=True)
    w64 = (torch.randn(K, Nout, dtype=torch.float64) / K ** 0.5).requires_grad_(True)
    b64 = torch.randn(Nout // div, dtype=torch.float64, requires_grad=True)
    gw = torch.randn(N, Nout, dtype=torch.float64)
    out64 = O.action_net_forward(ang64, it64, L, tr) @ w64 + b64.repeat_interleave(div)
    (out64 * gw).sum().backward()
    a, it, w, b = (t.detach().float().cuda().requires_grad_(True) for t in (ang64, it64, w64, b64))
    out = _ops.ActionGemm.apply(a, it, w, b, div, L, tr, chunk)
    (out * gw.float().cuda()).sum().backward()
    assert rel_rms(out, out64.detach()) < TF32_TOL
    for got, want, what, tol in ((a.grad, ang64.grad, "g_angles", TF32_TOL), (it.grad, it64.grad, "g_item_rep", TF32_TOL),
                                 (w.grad, w64.grad, "g_weight", TF32_TOL if wgrad_tf32 else 1e-4), (b.grad, b64.grad, "g_bias", 1e-4)):
        assert rel_rms(got, want) < tol, what
    assert torch.backends.cuda.matmul.allow_tf32 == allow_before          # the global switch is restored


@pytest.mark.parametrize("kind,L,hidden,N", [("deconv", 8, 50, 1000), ("deconv", 6, 200, 1024), ("mlp", 4, 0, 777)])
def test_action_net_fused_consumer_module(mods, kind, L, hidden, N):
    """ActionNet(fuse_consumer=True) in front of a DeconvNet-shaped stack / with the MLP: same forward as the unfused module
    and as float64 to TF32 accuracy.  (The unfused module's own first layer is TF32 too: cuDNN's default for FP32 convs.)"""
    _ops, dc = mods
    torch.manual_seed(L + hidden)
    C = 10
    Mh = (L + 1) ** 2

    class View(torch.nn.Module):                      # experiments/utils.py:36-42
        def __init__(self, *v):
            super().__init__()
            self.v = v

        def forward(self, x):
            return x.view(*self.v)
    if kind == "deconv":
        deconv = torch.nn.Sequential(View(-1, Mh * C, 1, 1), torch.nn.ConvTranspose2d(Mh * C, hidden, 4, 1, 0), torch.nn.ReLU())
        net = dc.ActionNet(L, deconv, rep_copies=C).cuda()
    else:
        net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, with_mlp=True).cuda()
    ang64 = O.group_matrix_to_eazyz(O.random_group_matrices(N, dtype=torch.float64))
    a = ang64.float().cuda()

    def run(fuse):
        net.fuse_consumer = fuse
        net.zero_grad(set_to_none=True)
        x = a.clone().requires_grad_(True)
        out = net(x)
        w = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        (out * w).sum().backward()
        return out.detach(), x.grad, {k: p.grad.clone() for k, p in net.named_parameters()}
    out_f, ga_f, gp_f = run(True)
    out_u, ga_u, gp_u = run(False)
    import copy
    y = O.action_net_forward(ang64, net.item_rep.detach().double().cpu(), L)
    out64 = copy.deepcopy(net.mlp if kind == "mlp" else net.deconv).cpu().double()(y).detach()
    assert out_f.shape == out_u.shape == tuple(out64.shape)
    assert rel_rms(out_f, out64) < TF32_TOL
    assert rel_rms(out_u, out64) < TF32_TOL
    # the ReLUs behind the layer make single gradients sensitive to which pre-activations TF32 rounding pushes across zero:
    # the two paths must agree on the bulk of the elements and in norm
    assert set(gp_f) == set(gp_u)
    for g1, g2, what in [(ga_f, ga_u, "g_angles")] + [(gp_f[k], gp_u[k], k) for k in gp_u]:
        scale = g2.double().pow(2).mean().sqrt().clamp_min(1e-30)
        bad = ((g1 - g2).abs().double() > 0.05 * scale).double().mean().item()
        assert bad < 0.02, (what, bad)
        assert abs(float(g1.double().norm() / g2.double().norm()) - 1.0) < 0.02, what
