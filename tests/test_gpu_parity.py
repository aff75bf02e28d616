"""GPU parity: CUDA kernels (through the C ABI / autograd wrappers) vs the FP64 oracle and the
reference-generated golden fixtures.  Run on the B200 box with ``-m gpu``.

Tolerance (north_star): rtol 1e-5 (+ atol 1e-5) in FP32 against an FP64 evaluation.  For the two
quantities where the reference's own FP32 code does not meet that bound against its FP64 self
(log_q and d/dsigma, SURVEY.md App. C: cancellation in 2-2cos and clamp-edge flips) the rule is
"at least as close to FP64 as the reference's FP32 arithmetic", checked by running the oracle in FP32.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import so3_oracle as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-5


@pytest.fixture(scope="module")
def lt():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import lie_vae_b200.lie_tools as m
    return m


def dev(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def close(actual, desired, rtol=RTOL, atol=ATOL, what=""):
    """allclose wherever the reference value is finite (the contract: match the reference where it is
    finite; at its singular points -- atan2(0,0), 0/0 at v = 0 -- the kernels may be better, not different)."""
    a = actual.detach().double().cpu().numpy() if torch.is_tensor(actual) else np.asarray(actual, dtype=np.float64)
    d = desired.detach().double().cpu().numpy() if torch.is_tensor(desired) else np.asarray(desired, dtype=np.float64)
    a, d = np.broadcast_arrays(a, d)
    ok = np.isfinite(d)
    np.testing.assert_allclose(a[ok], d[ok], rtol=rtol, atol=atol, err_msg=what)


def frac_outside(a, d, rtol=RTOL, atol=ATOL):
    a, d = np.asarray(a, dtype=np.float64), np.asarray(d, dtype=np.float64)
    return float(np.mean(np.abs(a - d) > atol + rtol * np.abs(d)))


def as_good_as_ref32(ours, ref64, ref32, what, floor=1.0):
    """ours (fp32 kernel) must be at least as close to the fp64 truth as the reference's own fp32 arithmetic
    (the oracle run in fp32).  Tolerances are relative to the tensor's scale (atol = 1e-5 * max(1, rms)), and
    ill-conditioned elements (atan2 near its cut, acos near the clamp, cancelling sums) may differ in *which*
    elements exceed the tolerance, so the comparison is on the fraction of outliers (<= 1.25x the reference's, +10 ppm
    or one element of slack) and on the worst error (<= 2x the reference's worst, or ``floor`` x the north-star atol)."""
    ours = ours.detach().double().cpu().numpy()
    ref64 = ref64.detach().double().cpu().numpy()
    ref32 = ref32.detach().double().cpu().numpy()
    ok = np.isfinite(ref64) & np.isfinite(ref32)        # singular points of the reference are exempt
    ours, ref64, ref32 = ours[ok], ref64[ok], ref32[ok]
    scale = max(1.0, float(np.sqrt(np.mean(ref64 ** 2)))) if ref64.size else 1.0
    atol = ATOL * scale
    f_ours, f_ref = frac_outside(ours, ref64, RTOL, atol), frac_outside(ref32, ref64, RTOL, atol)
    e_ours, e_ref = np.abs(ours - ref64).max(), np.abs(ref32 - ref64).max()
    slack = max(1e-5, 1.0 / max(ours.size, 1))
    assert f_ours <= 1.25 * f_ref + slack, "%s: %.3g of elements outside tol (reference fp32: %.3g)" % (what, f_ours, f_ref)
    assert e_ours <= max(2 * e_ref, floor * atol), "%s: max abs err %.3g (reference fp32: %.3g)" % (what, e_ours, e_ref)


def run_fn(fn, inputs, w, dtype):
    leaves = [dev(x, dtype).requires_grad_(True) for x in inputs]
    out = fn(*leaves)
    (out * dev(w, dtype)).sum().backward()
    return out, [l.grad for l in leaves]


# ----------------------------------------------------------------------- elementwise family vs golden
ROW_CASES = [
    ("rodrigues", "rodrigues", ["v"], ["gv"]),
    ("quat_to_mat", "quaternions_to_group_matrix", ["q"], ["gq"]),
    ("mat_to_quat", "group_matrix_to_quaternions", ["R"], ["gR"]),
    ("quat_to_eazyz", "quaternions_to_eazyz", ["q"], ["gq"]),
    ("mat_to_eazyz", "group_matrix_to_eazyz", ["R"], ["gR"]),
    ("s2s1", "s2s1rodrigues", ["s2", "s1"], ["gs2", "gs1"]),
    ("s2s2", "s2s2_gram_schmidt", ["v1", "v2"], ["gv1", "gv2"]),
    ("vector_to_eazyz", "vector_to_eazyz", ["v"], ["gv"]),
    ("log_map", "log_map", ["R"], ["gR"]),
]


@pytest.mark.parametrize("fixture,fname,ins,gs", ROW_CASES)
def test_rowop_f64_matches_reference(lt, fixture, fname, ins, gs):
    g = load_golden(fixture)
    out, grads = run_fn(getattr(lt, fname), [g[k] for k in ins], g["w"], torch.float64)
    close(out, g["out"], 1e-9, 1e-10, fname)
    for gr, k in zip(grads, gs):
        close(gr, g[k], 1e-8, 1e-9, fname + " grad " + k)


@pytest.mark.parametrize("fixture,fname,ins,gs", ROW_CASES)
def test_rowop_f32_matches_reference(lt, fixture, fname, ins, gs):
    g = load_golden(fixture)
    out, grads = run_fn(getattr(lt, fname), [g[k] for k in ins], g["w"], torch.float32)
    # fp32 reference arithmetic for the "as good as" rule
    leaves = [torch.tensor(g[k], dtype=torch.float32).requires_grad_(True) for k in ins]
    o32 = getattr(O, fname)(*leaves)
    (o32 * torch.tensor(g["w"], dtype=torch.float32)).sum().backward()
    as_good_as_ref32(out, torch.tensor(g["out"]), o32, fname)
    for gr, k, leaf in zip(grads, gs, leaves):
        as_good_as_ref32(gr, torch.tensor(g[k]), leaf.grad, fname + " grad " + k)


def test_algebra_maps(lt):
    g = load_golden("algebra")
    for dt in (torch.float32, torch.float64):
        v = dev(g["v"], dt).requires_grad_(True)
        X = lt.map_to_lie_algebra(v)
        close(X, g["hat"], 1e-7, 1e-7)
        close(lt.map_to_lie_vector(X), g["vee"], 1e-7, 1e-7)
        w = torch.randn_like(X)
        (X * w).sum().backward()
        vo = torch.tensor(g["v"]).requires_grad_(True)
        (O.map_to_lie_algebra(vo) * w.double().cpu()).sum().backward()
        close(v.grad, vo.grad, 1e-6, 1e-6)
        Xl = dev(g["hat"], dt).requires_grad_(True)
        wv = torch.randn(16, 3, device="cuda", dtype=dt)
        (lt.map_to_lie_vector(Xl) * wv).sum().backward()
        Xo = torch.tensor(g["hat"]).requires_grad_(True)
        (O.map_to_lie_vector(Xo) * wv.double().cpu()).sum().backward()
        close(Xl.grad, Xo.grad, 1e-6, 1e-6)


def test_known_answers(lt):
    R = lt.rodrigues(dev([[0.1, 0.2, 0.3]], torch.float64))
    close(R[0], [[0.9357548033, -0.2831649606, 0.2101917060], [0.3029327134, 0.9505806179, -0.0680313164],
                 [-0.1805400767, 0.1273345749, 0.9752903090]], 0, 1e-9)
    close(lt.map_to_lie_vector(lt.log_map(R))[0], [0.1, 0.2, 0.3], 0, 1e-12)
    M = lt.quaternions_to_group_matrix(dev([[0.1, 0.2, 0.3, 0.4]], torch.float64))
    close(M[0], [[2 / 15, 14 / 15, -1 / 3], [-2 / 3, 1 / 3, 2 / 3], [11 / 15, 2 / 15, 2 / 3]], 0, 1e-12)
    q = lt.group_matrix_to_quaternions(M)
    close(q[0], [0.1825741430, 0.3651482861, 0.5477224291, 0.7302969145], 0, 1e-9)
    close(lt.quaternions_to_eazyz(q)[0], [0.1798532748, 0.8410684190, 1.1071484928], 0, 1e-9)
    close(lt.group_matrix_to_eazyz(R)[0], [2.5273197923, 0.2227639061, -2.8285702969], 0, 1e-9)
    assert torch.equal(lt.rodrigues(torch.zeros(2, 3, device="cuda")), torch.eye(3, device="cuda").expand(2, 3, 3))


# ----------------------------------------------------------------------- the reference's own property tests
def test_log_exp_roundtrip(lt):
    # lie_tools.py:281-291, scales 0.1 and 10 at 1e-6 in float64
    torch.manual_seed(0)
    for scale in (0.1, 10.0):
        v0 = torch.randn(50, 3, dtype=torch.float64, device="cuda") * scale
        R = lt.rodrigues(v0)
        v = lt.map_to_lie_vector(lt.log_map(R))
        Rp = lt.rodrigues(v)
        vp = lt.map_to_lie_vector(lt.log_map(Rp))
        close(Rp, R, 1e-6, 1e-6)
        close(vp, v, 1e-6, 1e-6)


def test_orthogonality_and_det(lt):
    # lie_tools.py:396-425 at 1e-5 (fp32) / 1e-6 (fp64)
    torch.manual_seed(0)
    n = 10000
    s2 = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=-1)
    s1 = torch.nn.functional.normalize(torch.randn(n, 2, device="cuda"), dim=-1)
    R = lt.s2s1rodrigues(s2, s1)
    eye = torch.eye(3, device="cuda").expand(n, 3, 3)
    close(R @ R.transpose(1, 2), eye, 1e-5, 1e-5)
    close(torch.linalg.det(R), torch.ones(n), 1e-5, 1e-5)
    v1, v2 = torch.rand(2, n, 3, dtype=torch.float64, device="cuda")
    R = lt.s2s2_gram_schmidt(v1, v2)
    close(R @ R.transpose(1, 2), eye.double(), 1e-6, 1e-6)
    close(torch.linalg.det(R), torch.ones(n), 1e-6, 1e-6)
    q = torch.randn(100000, 4, dtype=torch.float64, device="cuda")
    R = lt.quaternions_to_group_matrix(q)
    close(R @ R.transpose(1, 2), torch.eye(3).expand(100000, 3, 3), 0, 1e-6)
    # mat -> quat -> mat round trip (lie_tools.py:301-304)
    r = lt.random_group_matrices(10000, dtype=torch.float64, device="cuda")
    close(lt.quaternions_to_group_matrix(lt.group_matrix_to_quaternions(r)), r, 1e-6, 1e-6)


def test_eazyz_vs_scipy(lt):
    # replaces the lie_learn KAT of lie_tools.py:294-320: R^T = Rz(a) Ry(b) Rz(c)  
    from scipy.spatial.transform import Rotation
    r = lt.random_group_matrices(2000, dtype=torch.float64, device="cuda")
    ea = lt.group_matrix_to_eazyz(r).cpu().numpy()
    back = Rotation.from_euler("ZYZ", ea).as_matrix()
    # 1e-4: the reference's 1e-6 guards (sqrt(1e-6+.), acos clamp at 1-1e-6) bound how well its angles
    # reproduce the matrix near beta = 0 / pi, in any precision
    np.testing.assert_allclose(back, r.cpu().numpy().transpose(0, 2, 1), atol=1e-4)


# ----------------------------------------------------------------------- large seeded comparison vs the oracle
@pytest.mark.parametrize("fname,shapes", [("rodrigues", [(3,)]), ("group_matrix_to_eazyz", None),
                                          ("quaternions_to_group_matrix", [(4,)]), ("group_matrix_to_quaternions", None)])
def test_rowop_large_vs_oracle(lt, fname, shapes):
    torch.manual_seed(1)
    n = 200003      # not a multiple of the tile
    if shapes is None:
        x = O.random_group_matrices(n, dtype=torch.float64)
    else:
        x = torch.randn(n, *shapes[0], dtype=torch.float64) * 1.3
    xo = x.clone().requires_grad_(True)
    out_o = getattr(O, fname)(xo)
    w = torch.randn_like(out_o)
    (out_o * w).sum().backward()
    x32 = x.float().clone().requires_grad_(True)
    o32 = getattr(O, fname)(x32)
    (o32 * w.float()).sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    out_g = getattr(lt, fname)(xg)
    (out_g * w.float().cuda()).sum().backward()
    as_good_as_ref32(out_g, out_o, o32, fname)
    as_good_as_ref32(xg.grad, xo.grad, x32.grad, fname + " grad")


def test_unaligned_and_noncontiguous_inputs(lt):
    torch.manual_seed(2)
    base = torch.randn(1001, 4, device="cuda")
    v = base[1:, 1:]                      # non-contiguous, storage offset 5 floats
    close(lt.rodrigues(v), O.rodrigues(v.double().cpu()), 1e-5, 1e-5)
    flat = torch.randn(3 * 777 + 1, device="cuda")[1:].view(777, 3)     # contiguous but only 4-byte aligned
    close(lt.rodrigues(flat), O.rodrigues(flat.double().cpu()), 1e-5, 1e-5)
    assert lt.rodrigues(torch.empty(0, 3, device="cuda")).shape == (0, 3, 3)
    close(lt.rodrigues(torch.randn(2, 5, 3, device="cuda")).shape, (2, 5, 3, 3), 0, 0)


def test_errors(lt):
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        lt.rodrigues(torch.randn(4, 3))
    with pytest.raises(TypeError):
        lt.rodrigues(torch.randn(4, 3, device="cuda").half())
    with pytest.raises(AssertionError):
        lt.map_to_lie_algebra(torch.randn(4, 4, device="cuda"))
    with pytest.raises(AssertionError):
        lt.group_matrix_to_quaternions(torch.randn(4, 3, 2, device="cuda"))
    with pytest.raises(AssertionError):
        lt.quaternions_to_eazyz(torch.randn(4, 3, device="cuda"))
    with pytest.raises(AssertionError):
        lt.wigner_d_matrix(torch.randn(4, 2, device="cuda"), 1)
    with pytest.raises(NotImplementedError):
        lt.wigner_d_matrix(torch.randn(4, 3, device="cuda"), 33)


# ----------------------------------------------------------------------- utils.logsumexp / IWAE bound
def test_logsumexp_matches_reference(lt):
    import lie_vae_b200.utils as ut
    g = load_golden("logsumexp")
    for dtype, rtol, atol in ((torch.float64, 1e-12, 1e-12), (torch.float32, RTOL, ATOL)):
        for dim, tag in ((1, "dim1"), (0, "dim0")):
            x = torch.tensor(g["x"], dtype=dtype, device="cuda").requires_grad_(True)
            out = ut.logsumexp(x, dim)
            (out * torch.tensor(g["w_" + tag], dtype=dtype, device="cuda")).sum().backward()
            close(out, g["out_" + tag], rtol, atol, "lse " + tag)
            close(x.grad, g["gx_" + tag], rtol, atol, "g lse " + tag)
            assert ut.logsumexp(x, dim, keepdim=True).shape == torch.logsumexp(x, dim, keepdim=True).shape
        close(ut.logsumexp(torch.tensor(g["x"], dtype=dtype, device="cuda")), g["out_all"], rtol, atol, "lse all")


def test_logsumexp_edge_cases_and_sizes(lt):
    import lie_vae_b200.utils as ut
    inf = float("inf")
    x = torch.tensor([[-inf, 0.0, inf, 1.0, float("nan"), -inf, 1000.0],
                      [-inf, -inf, 3.0, 2.0, 0.0, 5.0, 1000.0],
                      [-inf, 1.0, 1.0, -1e4, 1.0, -inf, -1000.0]], device="cuda")
    got, ref = ut.logsumexp(x, 0), torch.logsumexp(x, 0)
    assert torch.equal(torch.isnan(got), torch.isnan(ref))
    ok = ~torch.isnan(ref)
    assert torch.allclose(got[ok], ref[ok], rtol=1e-6, atol=1e-6) and got[0] == -inf and got[2] == inf
    torch.manual_seed(0)
    for shape, dim in (((500, 1), 0), ((1, 77), 0), ((9, 1000, 5), 1), ((33, 70000), 0), ((4, 5, 6), -1)):
        x = (torch.randn(*shape, device="cuda", dtype=torch.float64) * 30).requires_grad_(True)
        got = ut.logsumexp(x, dim)
        got.sum().backward()
        x2 = x.detach().clone().requires_grad_(True)
        ref = torch.logsumexp(x2, dim)
        ref.sum().backward()
        assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12) and torch.allclose(x.grad, x2.grad, rtol=1e-11, atol=1e-13)
    with pytest.raises(RuntimeError):
        ut.logsumexp(torch.zeros(3, 3), 0)


def test_iwae_log_likelihood_path(lt):
    """VAE.log_likelihood (vae.py:164-171; main.py:134-143 uses n = 500, B = 1): reparameterize with n samples of one
    datapoint -> action decoder -> per-sample reconstruction term -> importance-weighted bound, against the oracle."""
    import lie_vae_b200.reparameterize as rp
    import lie_vae_b200.decoders as dc
    import lie_vae_b200.utils as ut
    torch.manual_seed(1)
    n, B, L, C, k = 500, 1, 4, 3, 10
    M = (L + 1) ** 2
    mu64 = O.random_group_matrices(B, dtype=torch.float64)
    sg64 = torch.rand(B, 3, dtype=torch.float64) * 0.5 + 0.1
    eps64 = torch.randn(n, B, 3, dtype=torch.float64)
    item64, x64 = torch.randn(M, C, dtype=torch.float64), torch.randn(B, M * C, dtype=torch.float64)
    z, lq = O.so3_reparameterize(mu64, sg64, eps64, k)
    rec = O.action_net_forward(O.group_matrix_to_eazyz(z.reshape(-1, 3, 3)), item64, L).reshape(n, B, -1)
    w = -((rec - x64) ** 2).sum(-1) + O.so3_log_prior(z) - lq
    ref = (O.logsumexp(w, 0) - math.log(n)).mean()
    zc, lqc = rp.so3_reparameterize(mu64.float().cuda(), sg64.float().cuda(), eps64.float().cuda(), k)
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C).cuda()
    net.item_rep.data = item64.float().cuda()
    recc = net(lt.group_matrix_to_eazyz(zc.view(-1, 3, 3))).reshape(n, B, -1)
    log_p_x_z = -((recc - x64.float().cuda()) ** 2).sum(-1)
    log_p_z = torch.full((n, B), rp.LOG_PRIOR_SO3, dtype=torch.float64, device="cuda")
    got = ut.iwae_log_likelihood(log_p_x_z, log_p_z, lqc)
    assert abs(got.item() - ref.item()) < 5e-4 * max(1.0, abs(ref.item())), (got.item(), ref.item())


@pytest.mark.parametrize("n,B,L,C", [(500, 1, 8, 10), (500, 1, 4, 3), (7, 33, 6, 10), (3, 1000, 2, 1)])
def test_iwae_fused_reconstruction(lt, n, B, L, C):
    """The reconstruction term of VAE.log_likelihood fused into the Wigner forward (lv_wigner_recon_sse_f32): per-sample
    sum_{m,c} (D item_rep - x[i % B])^2 against the float64 oracle, and ``utils.action_log_likelihood`` (sampler -> Euler ->
    fused action + SSE -> log-sum-exp over n) against the oracle's IWAE bound on the same noise."""
    import lie_vae_b200.reparameterize as rp
    import lie_vae_b200.decoders as dc
    import lie_vae_b200.utils as ut
    from lie_vae_b200 import _ops
    torch.manual_seed(n + B + L)
    M = (L + 1) ** 2
    ang64 = O.group_matrix_to_eazyz(O.random_group_matrices(n * B, dtype=torch.float64))
    item64, x64 = torch.randn(M, C, dtype=torch.float64), torch.randn(B, M, C, dtype=torch.float64)
    y64 = O.action_net_forward(ang64, item64, L).reshape(n, B, M, C)
    want = ((y64 - x64) ** 2).sum((-1, -2)).reshape(-1)
    got = _ops.wigner_recon_sse(ang64.float().cuda(), item64.float().cuda(), x64.float().cuda(), L)
    close(got, want, 2e-5, 2e-5 * float(want.abs().max()), "sse")
    # the whole evaluation path from encoder features, fixed noise
    k = 10
    rep = rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.AlgebraMean(10), k=k).cuda()
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C).cuda()
    net.item_rep.data = item64.float().cuda()
    feats = torch.randn(B, 10, device="cuda")
    eps = torch.randn(n, B, 3, device="cuda")
    rep.reparameterize.sample_noise = lambda n_=1, like=None: eps
    got_ll = ut.action_log_likelihood(rep, net, feats, x64.float().cuda(), n)
    with torch.no_grad():
        mu64 = O.rodrigues(rep.mean_module.map(feats).double().cpu())
        sg64 = torch.nn.functional.softplus(rep.reparameterize.sigma_linear(feats)).double().cpu()
    z, lq = O.so3_reparameterize(mu64, sg64, eps.double().cpu(), k)
    rec = O.action_net_forward(O.group_matrix_to_eazyz(z.reshape(-1, 3, 3)), item64, L).reshape(n, B, M, C)
    w = -((rec - x64) ** 2).sum((-1, -2)) + O.so3_log_prior(z) - lq
    ref = (O.logsumexp(w, 0) - math.log(n)).mean()
    assert abs(got_ll.item() - ref.item()) < 1e-3 * max(1.0, abs(ref.item())), (got_ll.item(), ref.item())
