"""GPU parity for the two hot kernels: fused SO(3) reparameterize(+log-density) and the Wigner-D action.

CUDA (through the C ABI) vs golden fixtures from the unmodified reference, vs the FP64 oracle on
seeded inputs, and size-independent properties at BASELINE.json's full sizes.  ``-m gpu``.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import so3_oracle as O
from test_gpu_parity import RTOL, ATOL, as_good_as_ref32, close, dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import lie_vae_b200.lie_tools as lt
    import lie_vae_b200.reparameterize as rp
    import lie_vae_b200.decoders as dc
    return lt, rp, dc


# ----------------------------------------------------------------------- SO(3) reparameterize
def oracle_reparam(mu, sigma, eps, k, wz, wl, dtype):
    mu = torch.tensor(mu, dtype=dtype).requires_grad_(True)
    sg = torch.tensor(sigma, dtype=dtype).requires_grad_(True)
    z, lq = O.so3_reparameterize(mu, sg, torch.tensor(eps, dtype=dtype), k)
    ((z * torch.tensor(wz, dtype=dtype)).sum() + (lq * torch.tensor(wl, dtype=dtype)).sum()).backward()
    return z.detach(), lq.detach(), mu.grad, sg.grad


@pytest.mark.parametrize("name", ["so3_reparam_k3", "so3_reparam_k10", "so3_reparam_n5", "so3_reparam_iwae",
                                  "so3_reparam_ka6"])
def test_reparam_matches_reference(mods, name):
    _, rp, _ = mods
    g = load_golden(name)
    k = int(g["k"])
    mu, sg = dev(g["mu"]).requires_grad_(True), dev(g["sigma"]).requires_grad_(True)
    z, lq = rp.so3_reparameterize(mu, sg, dev(g["eps"]), k)
    ((z * dev(g["wz"])).sum() + (lq * dev(g["wl"])).sum()).backward()
    z32, lq32, gmu32, gsg32 = oracle_reparam(g["mu"], g["sigma"], g["eps"], k, g["wz"], g["wl"], torch.float32)
    close(z, g["z"], RTOL, ATOL, "z")
    close(mu.grad, g["gmu"], RTOL, ATOL, "g_mu")
    as_good_as_ref32(lq, torch.tensor(g["log_q"]), lq32, "log_q")
    as_good_as_ref32(sg.grad, torch.tensor(g["gsigma"]), gsg32, "g_sigma")


@pytest.mark.parametrize("name", ["so3_reparam_k3", "so3_reparam_k10", "so3_reparam_n5", "so3_reparam_iwae",
                                  "so3_reparam_ka6"])
def test_reparam_f64_matches_reference(mods, name):
    """The float64 instantiation of the fused kernels against the reference's float64 fixtures (values and
    gradients), directly and through the Euler-fused variant."""
    lt, rp, _ = mods
    g = load_golden(name)
    k = int(g["k"])
    d64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64, device="cuda")   # noqa: E731
    tol = dict(rtol=1e-8, atol=1e-9)
    mu, sg = d64(g["mu"]).requires_grad_(True), d64(g["sigma"]).requires_grad_(True)
    z, lq = rp.so3_reparameterize(mu, sg, d64(g["eps"]), k)
    assert z.dtype == torch.float64 and lq.dtype == torch.float64
    ((z * d64(g["wz"])).sum() + (lq * d64(g["wl"])).sum()).backward()
    for got, ref, what in ((z, g["z"], "z"), (lq, g["log_q"], "log_q"), (mu.grad, g["gmu"], "g_mu"), (sg.grad, g["gsigma"], "g_sigma")):
        np.testing.assert_allclose(got.detach().cpu().numpy(), ref, err_msg=what, **tol)
    # Euler-fused variant == group_matrix_to_eazyz of the plain one, gradients included
    wa = torch.randn(z.shape[:-2] + (3,), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    mu1, sg1 = d64(g["mu"]).requires_grad_(True), d64(g["sigma"]).requires_grad_(True)
    ang1, lq1 = rp.so3_reparameterize_eazyz(mu1, sg1, d64(g["eps"]), k)
    ((ang1 * wa).sum() + (lq1 * d64(g["wl"])).sum()).backward()
    mu2, sg2 = d64(g["mu"]).requires_grad_(True), d64(g["sigma"]).requires_grad_(True)
    z2, lq2 = rp.so3_reparameterize(mu2, sg2, d64(g["eps"]), k)
    ang2 = lt.group_matrix_to_eazyz(z2)
    ((ang2 * wa).sum() + (lq2 * d64(g["wl"])).sum()).backward()
    for a, b, what in ((ang1, ang2, "angles"), (mu1.grad, mu2.grad, "g_mu"), (sg1.grad, sg2.grad, "g_sigma")):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=1e-9, atol=1e-9, err_msg=what)


@pytest.mark.parametrize("k", [3, 10, 1, 5, 0])
def test_reparam_large_vs_oracle(mods, k):
    _, rp, _ = mods
    torch.manual_seed(0)
    B = 100003
    mu = O.random_group_matrices(B, dtype=torch.float64)
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, dtype=torch.float64))
    sigma[: B // 4] = 0.02 + 2.48 * torch.rand(B // 4, 3, dtype=torch.float64)     # stress set U(0.02, 2.5)
    eps = torch.randn(1, B, 3, dtype=torch.float64)
    wz, wl = torch.randn(1, B, 3, 3, dtype=torch.float64), torch.randn(1, B, dtype=torch.float64)
    z64, lq64, gmu64, gsg64 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float64)
    z32, lq32, gmu32, gsg32 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float32)
    mug, sgg = mu.float().cuda().requires_grad_(True), sigma.float().cuda().requires_grad_(True)
    z, lq = rp.so3_reparameterize(mug, sgg, eps.float().cuda(), k)
    ((z * wz.float().cuda()).sum() + (lq * wl.float().cuda()).sum()).backward()
    close(z, z64, RTOL, ATOL, "z")
    as_good_as_ref32(mug.grad, gmu64, gmu32, "g_mu")
    as_good_as_ref32(lq, lq64, lq32, "log_q")
    as_good_as_ref32(sgg.grad, gsg64, gsg32, "g_sigma")


@pytest.mark.parametrize("k", [0, 1, 3, 10])
@pytest.mark.parametrize("regime", ["narrow", "wide", "antipodal"])
def test_reparam_winding_regimes(mods, k, regime):
    # the winding sum is evaluated relative to the nearest winding (reparam_core.cuh): exercise the cases where that choice
    # matters -- sharply peaked densities (a x ~ 1e4), angles beyond the last winding (the reference point is clamped to
    # +-K), and angles next to pi where two windings carry the same weight
    _, rp, _ = mods
    torch.manual_seed(11)
    B = 20011
    mu = O.random_group_matrices(B, dtype=torch.float64)
    eps = torch.randn(1, B, 3, dtype=torch.float64)
    if regime == "narrow":
        sigma = 0.005 + 0.045 * torch.rand(B, 3, dtype=torch.float64)
    elif regime == "wide":
        sigma = 3.0 + 17.0 * torch.rand(B, 3, dtype=torch.float64)            # theta up to ~ 100 > (2K+1) pi
    else:
        sigma = 0.3 + 2.0 * torch.rand(B, 3, dtype=torch.float64)
        v = eps[0] * sigma
        target = math.pi * (1 + 2 * torch.randint(0, 3, (B, 1)).double()) + 1e-3 * torch.randn(B, 1, dtype=torch.float64)
        eps = (v / v.norm(dim=-1, keepdim=True) * target / sigma)[None]       # |eps * sigma| = (2j+1) pi + O(1e-3)
    wz, wl = torch.randn(1, B, 3, 3, dtype=torch.float64), torch.randn(1, B, dtype=torch.float64)
    z64, lq64, gmu64, gsg64 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float64)
    _, lq32, gmu32, gsg32 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float32)
    mug, sgg = mu.float().cuda().requires_grad_(True), sigma.float().cuda().requires_grad_(True)
    z, lq = rp.so3_reparameterize(mug, sgg, eps.float().cuda(), k)
    ((z * wz.float().cuda()).sum() + (lq * wl.float().cuda()).sum()).backward()
    assert torch.isfinite(lq).all() and torch.isfinite(sgg.grad).all() and torch.isfinite(mug.grad).all()
    as_good_as_ref32(lq, lq64, lq32, "log_q")
    as_good_as_ref32(mug.grad, gmu64, gmu32, "g_mu")
    as_good_as_ref32(sgg.grad, gsg64, gsg32, "g_sigma")
    # float64 instantiation against the float64 oracle
    mud, sgd = mu.cuda().requires_grad_(True), sigma.cuda().requires_grad_(True)
    zd, lqd = rp.so3_reparameterize(mud, sgd, eps.cuda(), k)
    ((zd * wz.cuda()).sum() + (lqd * wl.cuda()).sum()).backward()
    scale = lambda t: max(1.0, float(t.abs().max()))                             # noqa: E731
    assert (lqd.cpu() - lq64).abs().max().item() < 1e-9 * scale(lq64)
    assert (sgd.grad.cpu() - gsg64).abs().max().item() < 1e-8 * scale(gsg64)
    assert (mud.grad.cpu() - gmu64).abs().max().item() < 1e-8 * scale(gmu64)


def test_reparam_multisample_broadcast(mods):
    # n > 1 with B not a multiple of anything: tiles wrap over the broadcast mu / sigma rows
    _, rp, _ = mods
    torch.manual_seed(3)
    n, B, k = 7, 333, 3
    mu = O.random_group_matrices(B, dtype=torch.float64)
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, dtype=torch.float64))
    eps = torch.randn(n, B, 3, dtype=torch.float64)
    wz, wl = torch.randn(n, B, 3, 3, dtype=torch.float64), torch.randn(n, B, dtype=torch.float64)
    z64, lq64, gmu64, gsg64 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float64)
    _, lq32, gmu32, gsg32 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float32)
    mug, sgg = mu.float().cuda().requires_grad_(True), sigma.float().cuda().requires_grad_(True)
    z, lq = rp.so3_reparameterize(mug, sgg, eps.float().cuda(), k)
    assert z.shape == (n, B, 3, 3) and lq.shape == (n, B)
    ((z * wz.float().cuda()).sum() + (lq * wl.float().cuda()).sum()).backward()
    close(z, z64, RTOL, ATOL)
    as_good_as_ref32(mug.grad, gmu64, gmu32, "g_mu")
    as_good_as_ref32(lq, lq64, lq32, "log_q")
    as_good_as_ref32(sgg.grad, gsg64, gsg32, "g_sigma")


@pytest.mark.parametrize("n,B", [(1, 1024), (3, 1024), (3, 1028), (3, 1030), (2, 256), (5, 260), (1, 255), (1, 257)])
@pytest.mark.parametrize("euler", [False, True])
def test_reparam_tile_paths(mods, n, B, euler):
    # full tiles move with TMA bulk copies, tiles that wrap over the broadcast rows / start on an unaligned broadcast row /
    # are ragged, and tensors that are not 16-byte aligned, take the cp.async path: all must agree with the oracle, and
    # an unaligned view of the same data must give bit-identical results
    lt, rp, _ = mods
    torch.manual_seed(5)
    k = 3
    mu = O.random_group_matrices(B, dtype=torch.float64)
    # small sigma keeps theta away from 2 pi, where half an ulp of theta moves g_sigma by 1e-3 (1 / (2 - 2 cos theta) next
    # to its clamp): this test is about tile handling, the ill-conditioned regimes are test_reparam_winding_regimes'
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, dtype=torch.float64) - 1.5)
    eps = torch.randn(n, B, 3, dtype=torch.float64)
    wz, wl = torch.randn(n, B, 3, 3, dtype=torch.float64), torch.randn(n, B, dtype=torch.float64)
    z64, lq64, gmu64, gsg64 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float64)
    _, lq32, gmu32, gsg32 = oracle_reparam(mu, sigma, eps, k, wz, wl, torch.float32)

    def shifted(t):          # same values, storage offset of one element: not 16-byte aligned
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        buf[1:] = t.reshape(-1)
        return buf[1:].view(t.shape)

    def run(unaligned):
        prep = shifted if unaligned else (lambda t: t)
        mug, sgg = prep(mu.float().cuda()).requires_grad_(True), prep(sigma.float().cuda()).requires_grad_(True)
        e = prep(eps.float().cuda())
        if euler:
            ang, lq = rp.so3_reparameterize_eazyz(mug, sgg, e, k)
            wa = torch.randn(ang.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
            ((ang * wa).sum() + (lq * wl.float().cuda()).sum()).backward()
            return ang, lq, mug.grad, sgg.grad
        z, lq = rp.so3_reparameterize(mug, sgg, e, k)
        ((z * prep(wz.float().cuda())).sum() + (lq * wl.float().cuda()).sum()).backward()
        return z, lq, mug.grad, sgg.grad

    out, lq, gmu, gsg = run(False)
    out_u, lq_u, gmu_u, gsg_u = run(True)
    for a, b in ((out, out_u), (lq, lq_u), (gmu, gmu_u), (gsg, gsg_u)):
        assert torch.equal(a, b)
    as_good_as_ref32(lq, lq64, lq32, "log_q")
    if not euler:
        close(out, z64, RTOL, ATOL, "z")
        as_good_as_ref32(gmu, gmu64, gmu32, "g_mu")
        as_good_as_ref32(gsg, gsg64, gsg32, "g_sigma")
    else:
        z, _ = rp.so3_reparameterize(mu.float().cuda(), sigma.float().cuda(), eps.float().cuda(), k)
        assert torch.equal(out, lt.group_matrix_to_eazyz(z))


@pytest.mark.parametrize("n,B", [(1, (1 << 19) + 77), (2, 1 << 18), (3, 256 * 700), (1, 1 << 20)])
@pytest.mark.parametrize("euler", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_reparam_pipelined_matches_tile_kernels(mods, n, B, euler, dtype):
    # launches with many full tiles run the persistent double-buffered kernels (+ a one-tile launch for the ragged tail);
    # an unaligned view of the same data is forced onto the one-tile-per-CTA cp.async kernels, which the tests above pin
    # to the oracle: values and gradients must agree to rounding
    _, rp, _ = mods
    if dtype == torch.float64 and B > (1 << 19):
        B //= 2
    g = torch.Generator("cuda").manual_seed(9)
    k = 3
    q = torch.randn(B, 4, device="cuda", dtype=dtype, generator=g)
    import lie_vae_b200.lie_tools as lt
    mu = lt.quaternions_to_group_matrix(q)
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, device="cuda", dtype=dtype, generator=g))
    eps = torch.randn(n, B, 3, device="cuda", dtype=dtype, generator=g)
    wl = torch.randn(n, B, device="cuda", dtype=dtype, generator=g)
    wo = torch.randn(n, B, 3 if euler else 9, device="cuda", dtype=dtype, generator=g)

    def shifted(t):
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        buf[1:] = t.reshape(-1)
        return buf[1:].view(t.shape)

    def run(unaligned):
        prep = shifted if unaligned else (lambda t: t)
        mug, sgg = prep(mu.clone()).requires_grad_(True), prep(sigma.clone()).requires_grad_(True)
        fn = rp.so3_reparameterize_eazyz if euler else rp.so3_reparameterize
        out, lq = fn(mug, sgg, prep(eps), k)
        ((out.reshape(n, B, -1) * prep(wo)).sum() + (lq * wl).sum()).backward()
        return out, lq, mug.grad, sgg.grad

    a, b = run(False), run(True)
    # same per-sample source, but two kernels: ptxas may contract a mul + add differently, so allow rounding-level
    # differences (amplified at ill-conditioned samples); a tile-handling bug shows up as O(1) errors
    tol = 1e-4 if dtype == torch.float32 else 1e-9
    for x, y, what in zip(a, b, ("out", "log_q", "g_mu", "g_sigma")):
        assert torch.isfinite(x).all(), what
        err = (x - y).abs().max().item()
        assert err <= tol * max(1.0, x.abs().max().item()), "%s: %.3g" % (what, err)


def test_reparam_full_size_properties(mods):
    # BASELINE config 2: B = 2^20.  z orthogonal with det 1; log_q finite; z^T mu^T = exp(-v)
    lt, rp, _ = mods
    torch.manual_seed(0)
    B = 1 << 20
    mu = lt.random_group_matrices(B, device="cuda")
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, device="cuda"))
    eps = torch.randn(1, B, 3, device="cuda")
    z, lq = rp.so3_reparameterize(mu, sigma, eps, 3)
    eye = torch.eye(3, device="cuda")
    assert (z[0] @ z[0].transpose(1, 2) - eye).abs().max().item() < 5e-6
    assert (torch.linalg.det(z[0]) - 1).abs().max().item() < 5e-6
    assert torch.isfinite(lq).all()
    # mu^T z = exp(v): recover v with the log map where it is well conditioned
    v = (eps * sigma)[0]
    theta = v.norm(dim=-1)
    ok = (theta > 0.1) & (theta < 3.0)
    rel = mu.transpose(1, 2) @ z[0]
    v_back = lt.map_to_lie_vector(lt.log_map(rel))
    assert (v_back[ok] - v[ok]).abs().max().item() < 2e-4
    # log_q is invariant to the mean
    _, lq2 = rp.so3_reparameterize(lt.random_group_matrices(B, device="cuda"), sigma, eps, 3)
    assert torch.equal(lq, lq2)


def test_module_api(mods):
    lt, rp, _ = mods
    torch.manual_seed(0)
    m = rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.AlgebraMean(10), k=3).cuda()
    x = torch.randn(64, 10, device="cuda")
    z = m(x, n=4)
    assert z.shape == (4, 64, 3, 3)
    lq, lp, kl = m.log_posterior(), m.log_prior(), m.kl()
    assert lq.shape == (4, 64) and lq.dtype == torch.float32
    assert lp.shape == (4, 64) and lp.dtype == torch.float64 and abs(lp[0, 0].item() + 4.368901313378636) < 1e-12
    assert kl.shape == (64,) and kl.dtype == torch.float64
    # same numbers as the oracle fed with the module's own mu / sigma / eps
    zo, lqo = O.so3_reparameterize(m.mu_lie.double().cpu(), m.reparameterize.sigma.double().cpu(),
                                   m.reparameterize.eps.double().cpu(), 3)
    close(z, zo, RTOL, ATOL)
    close(lq, lqo, 1e-4, 1e-4)
    close(m.v, m.reparameterize.eps * m.reparameterize.sigma, 0, 0)
    (kl.sum() + z.sum()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert sorted(m.state_dict().keys()) == ["mean_module.map.bias", "mean_module.map.weight",
                                            "reparameterize.sigma_linear.bias", "reparameterize.sigma_linear.weight"]
    for mean in (rp.QuaternionMean(10), rp.S2S1Mean(10), rp.S2S2Mean(10)):
        mm = rp.SO3reparameterize(rp.N0reparameterize(10, 3, fixed_sigma=0.3), mean.cuda(), k=10).cuda()
        zz = mm(x)
        assert (zz[0] @ zz[0].transpose(1, 2) - torch.eye(3, device="cuda")).abs().max().item() < 1e-5
        mm.kl().sum().backward()
    m.deterministic()
    assert torch.equal(m(x, n=2)[1], m.mu_lie)


# ----------------------------------------------------------------------- Wigner
def test_wigner_d_matches_reference(mods):
    lt, _, _ = mods
    g = load_golden("wigner_d")
    ang = dev(g["angles"])
    for l in range(9):
        close(lt.wigner_d_matrix(ang, l), g["D%d" % l], RTOL, ATOL, "D%d" % l)


@pytest.mark.parametrize("tag", ["L8C3", "L3C1", "L5C10"])
@pytest.mark.parametrize("tr", ["N", "T"])
def test_block_wigner_matches_reference(mods, tag, tr):
    lt, _, _ = mods
    g = load_golden("block_wigner_%s_%s" % (tag, tr))
    L = int(g["max_degree"])
    a, s = dev(g["angles"]).requires_grad_(True), dev(g["spectrum"]).requires_grad_(True)
    out = lt.block_wigner_matrix_multiply(a, s, L, transpose=(tr == "T"))
    (out * dev(g["w"])).sum().backward()
    close(out, g["out"], RTOL, ATOL)
    close(s.grad, g["gspectrum"], RTOL, ATOL)
    close(a.grad, g["gangles"], 2e-5, 2e-5)


@pytest.mark.parametrize("name", ["action_net_L8C10", "action_net_L3C3"])
def test_action_net_matches_reference(mods, name):
    _, _, dc = mods
    g = load_golden(name)
    L, tr = int(g["degrees"]), bool(int(g["transpose"]))
    C = g["item_rep"].shape[1]
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, transpose=tr).cuda()
    net.item_rep.data = dev(g["item_rep"])
    a = dev(g["angles"]).requires_grad_(True)
    out = net(a)
    (out * dev(g["w"])).sum().backward()
    close(out, g["out"], RTOL, ATOL)
    close(net.item_rep.grad, g["gitem"], RTOL, 2e-5)
    close(a.grad, g["gangles"], 2e-5, 5e-5)


@pytest.mark.parametrize("L,C,N,tr", [(8, 10, 4099, False), (8, 10, 1000, True), (6, 10, 2048, False), (3, 3, 777, False),
                                      (8, 1, 513, False), (2, 32, 300, True), (0, 4, 65, False)])
def test_action_net_vs_oracle(mods, L, C, N, tr):
    _, _, dc = mods
    torch.manual_seed(L * 100 + C)
    M = (L + 1) ** 2
    ang = O.group_matrix_to_eazyz(O.random_group_matrices(N, dtype=torch.float64))
    item = torch.randn(M, C, dtype=torch.float64)
    w = torch.randn(N, M * C, dtype=torch.float64)

    def run_oracle(dt):
        a, it = ang.detach().clone().to(dt).requires_grad_(True), item.detach().clone().to(dt).requires_grad_(True)
        out = O.action_net_forward(a, it, L, tr)
        (out * w.to(dt)).sum().backward()
        return out.detach(), a.grad, it.grad
    o64, ga64, gi64 = run_oracle(torch.float64)
    o32, ga32, gi32 = run_oracle(torch.float32)
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, transpose=tr).cuda()
    net.item_rep.data = item.float().cuda()
    a = ang.float().cuda().requires_grad_(True)
    out = net(a)
    (out * w.float().cuda()).sum().backward()
    as_good_as_ref32(out, o64, o32, "out")
    as_good_as_ref32(a.grad, ga64, ga32, "g_angles")
    # the batch-summed gradient: relative to its own scale (sum of N terms)
    scale = gi64.abs().max().item()
    assert (net.item_rep.grad.double().cpu() - gi64).abs().max().item() <= max((gi32.double() - gi64).abs().max().item(), 2e-6 * scale)


def test_wigner_reference_properties(mods):
    # lie_tools.py:337-357: orthogonality, W(g)W(g^-1) = I, W(b)W(a) = W(ab) (rtol/atol 1e-3 in the reference)
    lt, _, _ = mods
    torch.manual_seed(0)
    for l in range(9):
        r = lt.random_group_matrices(2000, device="cuda")
        D = lt.wigner_d_matrix(lt.group_matrix_to_eazyz(r), l)
        eye = torch.eye(2 * l + 1, device="cuda").expand_as(D)
        close(D @ D.transpose(1, 2), eye, 1e-4, 1e-5)
        Dinv = lt.wigner_d_matrix(lt.group_matrix_to_eazyz(r.transpose(1, 2)), l)
        close(D @ Dinv, eye, 1e-4, 2e-5)
        ra, rb = lt.random_group_matrices(2000, device="cuda"), lt.random_group_matrices(2000, device="cuda")
        wa = lt.wigner_d_matrix(lt.group_matrix_to_eazyz(ra), l)
        wb = lt.wigner_d_matrix(lt.group_matrix_to_eazyz(rb), l)
        wc = lt.wigner_d_matrix(lt.group_matrix_to_eazyz(ra.bmm(rb)), l)
        close(wb.bmm(wa), wc, 1e-3, 1e-3)
    # l = 1 block is the rotation itself in the (y,z,x) basis: D1(eazyz(R)) = P R^T P^T
    r = lt.random_group_matrices(100, device="cuda")
    D1 = lt.wigner_d_matrix(lt.group_matrix_to_eazyz(r), 1)
    P = torch.tensor([[0., 1, 0], [0, 0, 1], [1, 0, 0]], device="cuda")
    close(D1, P @ r.transpose(1, 2) @ P.t(), 1e-5, 1e-5)


def test_wigner_full_size_properties(mods):
    # BASELINE config 3: N = 65536, L = 8, C = 10.  D is orthogonal: per-degree norms are preserved;
    # D^T D = I through the transpose flag; linear in the spectrum; shared == expanded per-sample.
    lt, _, dc = mods
    torch.manual_seed(0)
    N, L, C = 65536, 8, 10
    M = (L + 1) ** 2
    ang = lt.group_matrix_to_eazyz(lt.random_group_matrices(N, device="cuda"))
    item = torch.randn(M, C, device="cuda")
    out = lt.block_wigner_matrix_multiply(ang, item.expand(N, -1, -1), L)
    assert out.shape == (N, M, C)
    start = 0
    for l in range(L + 1):
        d = 2 * l + 1
        n_in = item[start:start + d].pow(2).sum(0)
        n_out = out[:, start:start + d].pow(2).sum(1)
        assert ((n_out - n_in) / n_in).abs().max().item() < 2e-5
        start += d
    back = lt.block_wigner_matrix_multiply(ang, out, L, transpose=True)
    assert (back - item).abs().max().item() < 2e-5
    sub = slice(0, 4096)
    per_sample = lt.block_wigner_matrix_multiply(ang[sub], item.expand(4096, -1, -1).contiguous(), L)
    assert torch.equal(per_sample, out[sub])
    other = torch.randn(M, C, device="cuda")
    lin = lt.block_wigner_matrix_multiply(ang[sub], (2 * item - 3 * other).expand(4096, -1, -1), L)
    ref = 2 * out[sub] - 3 * lt.block_wigner_matrix_multiply(ang[sub], other.expand(4096, -1, -1), L)
    assert (lin - ref).abs().max().item() < 5e-5
    # backward: deterministic, and grad(item_rep) of sum(out * g) equals sum_n D_n^T g_n
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C).cuda()
    net.item_rep.data = item
    g = torch.randn(N, M * C, device="cuda")
    a = ang.clone().requires_grad_(True)
    (net(a) * g).sum().backward()
    g1 = net.item_rep.grad.clone()
    net.item_rep.grad = None
    a2 = ang.clone().requires_grad_(True)
    (net(a2) * g).sum().backward()
    assert torch.equal(g1, net.item_rep.grad) and torch.equal(a.grad, a2.grad)
    gt = lt.block_wigner_matrix_multiply(ang, g.view(N, M, C), L, transpose=True).double().sum(0)
    assert (g1.double() - gt).abs().max().item() < 1e-3 * gt.abs().max().item()


def test_wigner_errors(mods):
    lt, _, dc = mods
    a = torch.randn(4, 3, device="cuda")
    with pytest.raises(NotImplementedError):
        lt.block_wigner_matrix_multiply(a, torch.randn(4, 34 * 34, 2, device="cuda"), 33)
    with pytest.raises(ValueError):
        lt.block_wigner_matrix_multiply(a, torch.randn(4, 10, 2, device="cuda"), 2)
    with pytest.raises(AssertionError):
        dc.ActionNet(2, torch.nn.Sequential()).cuda()(torch.randn(4, 4, device="cuda"))
    assert lt.block_wigner_matrix_multiply(torch.empty(0, 3, device="cuda"), torch.empty(0, 9, 2, device="cuda"), 2).shape == (0, 9, 2)


@pytest.mark.parametrize("L", [8, 6])
@pytest.mark.parametrize("N", [2, 3, 11, 12, 13, 16, 17, 31, 48, 49, 12 * 148 + 2, 16 * 150 + 3, 16 * 600, 12 * 148 * 6 + 10])
@pytest.mark.parametrize("tr", [False, True])
def test_ws_backward_matches_cp_async_backward(mods, L, N, tr):
    """The degree-specialised TMA-fed backward (C = 10, degrees 0..8 / 0..6, 16-byte aligned g_y; 12-sample tiles, ragged last
    tile, odd last sample) against the cp.async kernel, which the same call falls back to when g_y is only 4-byte aligned;
    and against the oracle for the small sizes."""
    from lie_vae_b200 import _ops
    torch.manual_seed(N + L)
    M, C = (L + 1) ** 2, 10
    ang = O.group_matrix_to_eazyz(O.random_group_matrices(N, dtype=torch.float64)).float().cuda()
    item = torch.randn(M, C, device="cuda")
    g_aligned = torch.randn(N, M, C, device="cuda")
    g_unaligned = torch.empty(N * M * C + 1, device="cuda")[1:].view(N, M, C)
    g_unaligned.copy_(g_aligned)
    assert g_aligned.data_ptr() % 16 == 0 and g_unaligned.data_ptr() % 16 != 0

    def run(g):
        a, it = ang.clone().requires_grad_(True), item.clone().requires_grad_(True)
        out = _ops.WignerApply.apply(a, it, 0, L, tr)
        out.backward(g)
        return a.grad, it.grad
    ga1, gi1 = run(g_aligned)
    ga2, gi2 = run(g_unaligned)
    # angle gradients: body-frame generator formulation (warp-decoupled kernel) vs forward recompute (cp.async kernel)
    assert (ga1 - ga2).abs().max().item() <= 2e-5 * max(1.0, ga2.abs().max().item())
    scale = gi2.abs().max().item()
    assert (gi1 - gi2).abs().max().item() <= 2e-6 * scale          # only the summation order differs
    ga1b, gi1b = run(g_aligned)
    assert torch.equal(gi1, gi1b) and torch.equal(ga1, ga1b)        # run-to-run reproducible
    if N <= 64:
        a64, it64 = ang.double().cpu().requires_grad_(True), item.double().cpu().requires_grad_(True)
        (O.action_net_forward(a64, it64, L, tr) * g_aligned.double().cpu().view(N, -1)).sum().backward()
        close(gi1, it64.grad, 1e-5, 1e-5 * max(1.0, scale))
        close(ga1, a64.grad, 2e-5, 1e-4)


def test_ws_backward_stress_reproducible(mods):
    """Many back-to-back launches of the degree-specialised backward (persistent CTAs, math warps bound to degree groups,
    producer warps recycling the mbarrier-tracked tile ring): every launch must reproduce the first one bit for bit.
    Guards the full/empty hand-over protocol."""
    from lie_vae_b200 import _ops
    torch.manual_seed(5)
    N, L, C = 1 << 17, 8, 10
    M = (L + 1) ** 2
    ang = (torch.rand(N, 3, device="cuda") * 6.0 - 3.0)
    item = torch.randn(M, C, device="cuda")
    g = torch.randn(N, M, C, device="cuda")

    def run():
        a, it = ang.clone().requires_grad_(True), item.clone().requires_grad_(True)
        _ops.WignerApply.apply(a, it, 0, L, False).backward(g)
        return a.grad, it.grad
    ga0, gi0 = run()
    for _ in range(300):
        ga, gi = run()
        assert torch.equal(gi, gi0) and torch.equal(ga, ga0)
    torch.cuda.synchronize()


@pytest.mark.parametrize("k,n,B", [(3, 1, 100003), (10, 1, 4097), (5, 3, 333)])
def test_fused_reparam_eazyz(mods, k, n, B):
    """so3_reparameterize_eazyz == group_matrix_to_eazyz(so3_reparameterize(...)), forward and backward: against the
    two-kernel composition (bit-exact forward: same device code) and against the FP64 oracle."""
    lt, rp, _ = mods
    torch.manual_seed(k * 7 + n)
    mu = O.random_group_matrices(B, dtype=torch.float64)
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, dtype=torch.float64))
    eps = torch.randn(n, B, 3, dtype=torch.float64)
    wa, wl = torch.randn(n, B, 3, dtype=torch.float64), torch.randn(n, B, dtype=torch.float64)

    def oracle(dt):
        m, s = mu.to(dt).clone().requires_grad_(True), sigma.to(dt).clone().requires_grad_(True)
        z, lq = O.so3_reparameterize(m, s, eps.to(dt), k)
        ang = O.group_matrix_to_eazyz(z)
        ((ang * wa.to(dt)).sum() + (lq * wl.to(dt)).sum()).backward()
        return ang.detach(), lq.detach(), m.grad, s.grad
    a64, l64, gm64, gs64 = oracle(torch.float64)
    a32, l32, gm32, gs32 = oracle(torch.float32)

    def gpu(fused):
        m, s = mu.float().cuda().requires_grad_(True), sigma.float().cuda().requires_grad_(True)
        if fused:
            ang, lq = rp.so3_reparameterize_eazyz(m, s, eps.float().cuda(), k)
        else:
            z, lq = rp.so3_reparameterize(m, s, eps.float().cuda(), k)
            ang = lt.group_matrix_to_eazyz(z)
        ((ang * wa.float().cuda()).sum() + (lq * wl.float().cuda()).sum()).backward()
        return ang.detach(), lq.detach(), m.grad, s.grad
    af, lf, gmf, gsf = gpu(True)
    ac, lc, gmc, gsc = gpu(False)
    assert af.shape == (n, B, 3) and lf.shape == (n, B)
    assert torch.equal(af, ac) and torch.equal(lf, lc)
    # backward: same math, different kernels (FMA contraction differs); near gimbal lock the Euler pull-back is
    # ill-conditioned, so compare on the bulk of the elements
    for got, ref in ((gmf, gmc), (gsf, gsc)):
        bad = (got - ref).abs() > 1e-5 * ref.abs() + 1e-5 * max(1.0, ref.abs().max().item())
        assert bad.float().mean().item() < 1e-4
    as_good_as_ref32(af, a64, a32, "angles")
    as_good_as_ref32(lf, l64, l32, "log_q")
    as_good_as_ref32(gmf, gm64, gm32, "g_mu")
    as_good_as_ref32(gsf, gs64, gs32, "g_sigma")


def test_z_rot_mat(mods):
    lt, _, _ = mods
    ang = torch.tensor([0.3, -1.2, 2.5], device="cuda")
    for l in (0, 1, 4, 8):
        close(lt._z_rot_mat(ang, l), O.z_rot_mat(ang.double().cpu(), l), 1e-5, 2e-6)


@pytest.mark.parametrize("L,C,N", [(8, 10, 4112), (3, 3, 500)])
def test_wigner_backward_accumulate_flag(mods, L, C, N):
    """shared_spectrum = 3 adds the batch sum to gspectrum (TMA-fed and cp.async kernels): two micro-batches accumulated
    through the pipeline runner equal one call over the concatenated batch."""
    from lie_vae_b200.pipeline import FusedSO3ActionStep
    from lie_vae_b200 import _ops
    torch.manual_seed(9)
    M = (L + 1) ** 2
    ang = torch.rand(N, 3, device="cuda") * 6 - 3
    item = torch.randn(M, C, device="cuda")
    g = torch.randn(N, M, C, device="cuda")
    a, it = ang.clone().requires_grad_(True), item.clone().requires_grad_(True)
    _ops.WignerApply.apply(a, it, 0, L, False).backward(g)
    half = (N // 32) * 16
    step = FusedSO3ActionStep(N, N, L, C, 3, device="cuda")
    step.angles.copy_(ang)
    step.g_item.zero_()
    step.decode_backward(0, half, item, g[:half].contiguous(), accumulate=True)
    step.decode_backward(half, N, item, g[half:].contiguous(), accumulate=True)
    assert torch.equal(step.g_angles, a.grad)
    assert (step.g_item - it.grad).abs().max().item() <= 2e-6 * it.grad.abs().max().item()
    step.decode_backward(0, N, item, g, accumulate=False)          # overwrite mode ignores what was there
    assert (step.g_item - it.grad).abs().max().item() <= 2e-6 * it.grad.abs().max().item()


# ----------------------------------------------------------------------- encoder heads fused into the reparameterize kernel
@pytest.mark.parametrize("mode,n,B,Din,k", [("alg", 1, 1000, 10, 10), ("q", 3, 777, 10, 3), ("s2s2", 1, 513, 10, 10),
                                            ("alg", 2, 64, 32, 5), ("q", 1, 5, 7, 0), ("s2s2", 2, 300, 16, 3),
                                            ("s2s1", 1, 900, 10, 10), ("s2s1", 3, 77, 20, 3)])
def test_fused_heads_match_unfused_modules(mods, mode, n, B, Din, k):
    """SO3reparameterize with the heads inside the kernel (the default) against the same module with fuse_heads = False
    (Linear -> mean map kernel -> softplus -> reparameterize kernel): z, log_q, mu_lie, sigma, kl and the gradients of the
    features and of every head parameter."""
    _, rp, _ = mods
    torch.manual_seed(B + Din)
    mean_cls = {"alg": rp.AlgebraMean, "q": rp.QuaternionMean, "s2s2": rp.S2S2Mean, "s2s1": rp.S2S1Mean}[mode]
    mod = rp.SO3reparameterize(rp.N0reparameterize(Din, 3), mean_cls(Din), k=k).cuda()
    if mode == "s2s2":
        mod.mean_module.map.weight.data.uniform_(-1, 1)
    x0 = torch.randn(B, Din, device="cuda")
    eps = torch.randn(n, B, 3, device="cuda")
    wz, wl = torch.randn(n, B, 3, 3, device="cuda"), torch.randn(n, B, device="cuda")
    wm, ws = torch.randn(B, 3, 3, device="cuda"), torch.randn(B, 3, device="cuda")
    res = []
    for fuse in (True, False):
        mod.fuse_heads = fuse
        mod.zero_grad()
        mod.reparameterize.sample_noise = lambda n_=1, like=None: eps
        x = x0.clone().requires_grad_(True)
        z = mod(x, n)
        assert mod._fused == fuse
        # mu_lie and sigma are differentiable on both paths: regularisers / kl terms built from them reach the heads
        ((z * wz).sum() + (mod.log_posterior() * wl).sum() + mod.kl().sum().float()
         + (mod.mu_lie * wm).sum() + (mod.reparameterize.sigma * ws).sum() + mod.reparameterize.kl().sum()).backward()
        res.append((z.detach(), mod.log_posterior().detach(), mod.mu_lie.detach(), mod.reparameterize.sigma.detach(), x.grad,
                    {k_: p.grad.clone() for k_, p in mod.named_parameters()}))
    (z1, lq1, mu1, sg1, gx1, gp1), (z2, lq2, mu2, sg2, gx2, gp2) = res
    close(mu1, mu2, 1e-5, 2e-6, "mu")
    close(sg1, sg2, 1e-5, 1e-6, "sigma")
    close(z1, z2, 1e-5, 5e-6, "z")
    assert (lq1 - lq2).abs().max().item() <= 2e-4 * max(1.0, lq2.abs().max().item())
    for a, b, what in [(gx1, gx2, "g_x")] + [(gp1[k_], gp2[k_], k_) for k_ in gp1]:
        scale = max(1.0, b.abs().max().item())
        assert (a - b).abs().max().item() <= 2e-3 * scale, what


@pytest.mark.parametrize("mode", ["alg", "q", "s2s1", "s2s2"])
@pytest.mark.parametrize("euler", [True, False])
def test_fused_heads_functional_vs_oracle(mods, mode, euler):
    """so3_head_reparameterize (every mean map, pose as z or as Euler angles) against the float64 oracle composition
    (Linear -> mean map, softplus -> reparameterize [-> Euler]) including the gradients of features, weights and biases,
    with gradient also arriving through the mu / sigma outputs."""
    _, rp, _ = mods
    torch.manual_seed(9)
    B, Din, n, k = 700, 10, 2, 3
    dm = {"alg": 3, "q": 4, "s2s1": 5, "s2s2": 6}[mode]
    h64 = torch.randn(B, Din, dtype=torch.float64)
    W64, b64 = torch.randn(dm + 3, Din, dtype=torch.float64) * 0.5, torch.randn(dm + 3, dtype=torch.float64) * 0.5
    eps64 = torch.randn(n, B, 3, dtype=torch.float64)
    wp = torch.randn((n, B, 3) if euler else (n, B, 3, 3), dtype=torch.float64)
    wl, wm, ws = torch.randn(n, B, dtype=torch.float64), torch.randn(B, 3, 3, dtype=torch.float64), torch.randn(B, 3, dtype=torch.float64)

    def mean_map(pre):
        if mode == "alg":
            return O.rodrigues(pre)
        if mode == "q":
            return O.quaternions_to_group_matrix(pre)
        if mode == "s2s1":
            s2, s1 = pre[:, :3], pre[:, 3:]
            return O.s2s1rodrigues(s2 / s2.norm(p=2, dim=-1, keepdim=True), s1 / s1.norm(p=2, dim=-1, keepdim=True))
        v = pre.double().view(-1, 2, 3)                     # reparameterize.py:195-197: Gram-Schmidt in float64, cast back
        return O.s2s2_gram_schmidt(v[:, 0], v[:, 1]).to(pre.dtype)

    def run(dtype, dev):
        h, W, b = (t.detach().clone().to(dtype).to(dev).requires_grad_(True) for t in (h64, W64, b64))
        eps = eps64.to(dtype).to(dev)
        if dev == "cuda":
            pose, lq, mu, sg = rp.so3_head_reparameterize(h, W[:dm], b[:dm], W[dm:], b[dm:], eps, mode, k, euler=euler)
        else:
            pre = h @ W.t() + b
            mu, sg = mean_map(pre[:, :dm]), torch.nn.functional.softplus(pre[:, dm:])
            pose, lq = O.so3_reparameterize(mu, sg, eps, k)
            if euler:
                pose = O.group_matrix_to_eazyz(pose)
        c = lambda t: t.to(dtype).to(dev)
        ((pose * c(wp)).sum() + (lq * c(wl)).sum() + (mu * c(wm)).sum() + (sg * c(ws)).sum()).backward()
        return [t.detach().double().cpu() for t in (pose, lq, mu, sg, h.grad, W.grad, b.grad)]
    got, ref, ref32 = run(torch.float32, "cuda"), run(torch.float64, "cpu"), run(torch.float32, "cpu")
    names = ["pose", "log_q", "mu", "sigma", "g_h", "g_W", "g_b"]
    for nm, a, b, c in zip(names, got, ref, ref32):
        if nm in ("mu", "sigma"):
            close(a, b, 1e-5, 2e-6, nm)
        else:
            as_good_as_ref32(a, b, c, "%s/%s" % (mode, nm))


def test_fused_heads_nsample_and_fallbacks(mods):
    _, rp, _ = mods
    torch.manual_seed(2)
    mod = rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.AlgebraMean(10), k=3).cuda()
    x = torch.randn(40, 10, device="cuda", requires_grad=True)
    z0 = mod(x, 2)
    lq0, v0 = mod.log_posterior().detach().clone(), mod.v.detach().clone()
    z = mod.nsample(4)      # reparameterize.py:269-273: a pure function of the cached mu_lie / v -- same sample, nothing overwritten
    assert tuple(z.shape) == (2, 40, 3, 3)
    close(z, z0, 1e-5, 5e-6, "nsample")
    assert torch.equal(mod.log_posterior(), lq0) and torch.equal(mod.v, v0) and mod.z is z0
    z.sum().backward()      # mu_lie and sigma are differentiable outputs of the fused kernel: gradients reach the features
    assert x.grad is not None and x.grad.abs().sum().item() > 0 and mod.mean_module.map.weight.grad is not None
    assert mod.reparameterize.sigma_linear.weight.grad is not None
    # paths the fused kernel does not cover fall back to the separate launches: fixed sigma, wide features, float64
    for m2, xin in ((rp.SO3reparameterize(rp.N0reparameterize(10, 3, fixed_sigma=0.3), rp.AlgebraMean(10)).cuda(), torch.randn(8, 10, device="cuda")),
                    (rp.SO3reparameterize(rp.N0reparameterize(40, 3), rp.AlgebraMean(40)).cuda(), torch.randn(8, 40, device="cuda")),
                    (rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.QuaternionMean(10)).cuda().double(), torch.randn(8, 10, device="cuda", dtype=torch.float64))):
        out = m2(xin)
        assert not m2._fused and tuple(out.shape) == (1, 8, 3, 3) and tuple(m2.kl().shape) == (8,)
    mod.deterministic()
    assert torch.equal(mod(x)[0], mod.mu_lie)


# ----------------------------------------------------------------------- BASELINE sizes against the oracle (row subsets)
def test_reparam_baseline_size_subset_vs_oracle(mods):
    """BASELINE configs[1] (B = 2^20, k = 3) through the persistent kernels: a random 8 192-row subset of z, log_q, g_mu and
    g_sigma against the float64 oracle evaluated on exactly those rows (samples are independent)."""
    lt, rp, _ = mods
    torch.manual_seed(11)
    B, S, k = 1 << 20, 8192, 3
    mu = lt.random_group_matrices(B, device="cuda").requires_grad_(True)
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, device="cuda")).requires_grad_(True)
    eps = torch.randn(1, B, 3, device="cuda")
    wz, wl = torch.randn(1, B, 3, 3, device="cuda"), torch.randn(1, B, device="cuda")
    z, lq = rp.so3_reparameterize(mu, sigma, eps, k)
    ((z * wz).sum() + (lq * wl).sum()).backward()
    idx = torch.randperm(B, device="cuda")[:S]
    sub = lambda t, d=0: t.detach().index_select(d, idx).double().cpu()
    z64, lq64, gmu64, gsg64 = oracle_reparam(sub(mu), sub(sigma), sub(eps, 1), k, sub(wz, 1), sub(wl, 1), torch.float64)
    z32, lq32, gmu32, gsg32 = oracle_reparam(sub(mu), sub(sigma), sub(eps, 1), k, sub(wz, 1), sub(wl, 1), torch.float32)
    close(sub(z, 1), z64, RTOL, ATOL, "z")
    as_good_as_ref32(sub(mu.grad), gmu64, gmu32, "g_mu")
    as_good_as_ref32(sub(lq, 1), lq64, lq32, "log_q")
    as_good_as_ref32(sub(sigma.grad), gsg64, gsg32, "g_sigma")


def test_action_net_baseline_size_subset_vs_oracle(mods):
    """BASELINE configs[2] (N = 65 536, degrees 8, 10 channels) through ActionNet: a random 8 192-row subset of the output and
    of g_angles against the float64 oracle on those rows; g_item_rep (a sum over ALL rows) against the oracle's gradient for
    the subset plus the complement's contribution evaluated by the kernels' own transpose action in float64 accumulation."""
    lt, _, dc = mods
    torch.manual_seed(12)
    N, L, C, S = 65536, 8, 10, 8192
    M = (L + 1) ** 2
    ang = O.group_matrix_to_eazyz(O.random_group_matrices(N, dtype=torch.float64)).float().cuda()
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C).cuda()
    g = torch.randn(N, M * C, device="cuda")
    a = ang.clone().requires_grad_(True)
    out = net(a)
    (out * g).sum().backward()
    idx = torch.randperm(N, device="cuda")[:S]
    a64 = ang.index_select(0, idx).double().cpu().requires_grad_(True)
    it64 = net.item_rep.detach().double().cpu().requires_grad_(True)
    g_sub = g.index_select(0, idx).double().cpu()
    out64 = O.action_net_forward(a64, it64, L)
    (out64 * g_sub).sum().backward()
    a32 = a64.detach().float().requires_grad_(True)
    it32 = it64.detach().float().requires_grad_(True)
    out32 = O.action_net_forward(a32, it32, L)
    (out32 * g_sub.float()).sum().backward()
    as_good_as_ref32(out.detach().index_select(0, idx), out64.detach(), out32.detach(), "y")
    as_good_as_ref32(a.grad.index_select(0, idx), a64.grad, a32.grad, "g_angles")
    # g_item_rep: subset by the oracle + complement by sum_n D_n^T g_n (float64 accumulation of the float32 transpose action)
    mask = torch.ones(N, dtype=torch.bool, device="cuda")
    mask[idx] = False
    comp = lt.block_wigner_matrix_multiply(ang[mask], g[mask].view(-1, M, C), L, transpose=True).double().sum(0).cpu()
    want = it64.grad + comp
    got = net.item_rep.grad.double().cpu()
    assert (got - want).abs().max().item() <= 2e-5 * want.abs().max().item()


@pytest.mark.parametrize("variant", ["with_mlp", "buffer_item_rep", "buffer_transpose_mlp"])
def test_action_net_mlp_and_buffer_variants_vs_oracle(mods, variant):
    """ActionNet(with_mlp=True) (``decoders.py:39-41,58-59``) and the fixed ``item_rep`` buffer (``decoders.py:36-37``):
    forward and backward on the GPU against the float64 oracle action followed by the same MLP weights in float64."""
    _, _, dc = mods
    torch.manual_seed(21)
    N, L, C = 777, 4, 10
    M = (L + 1) ** 2
    tr = variant == "buffer_transpose_mlp"
    with_mlp = variant != "buffer_item_rep"
    fixed = None if variant == "with_mlp" else torch.randn(M, C)
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, with_mlp=with_mlp, item_rep=fixed, transpose=tr).cuda()
    assert ("item_rep" in dict(net.named_buffers())) == (fixed is not None)
    assert ("item_rep" in dict(net.named_parameters())) == (fixed is None)
    ang64 = O.group_matrix_to_eazyz(O.random_group_matrices(N, dtype=torch.float64))
    w = torch.randn(N, M * C, dtype=torch.float64)
    a = ang64.float().cuda().requires_grad_(True)
    out = net(a)
    (out * w.float().cuda()).sum().backward()

    def oracle(dtype):
        import copy
        a_ = ang64.detach().clone().to(dtype).requires_grad_(True)
        it = net.item_rep.detach().cpu().clone().to(dtype).requires_grad_(fixed is None)
        y = O.action_net_forward(a_, it, L, tr)
        mlp = copy.deepcopy(net.mlp).cpu().to(dtype) if with_mlp else None
        if mlp is not None:
            y = mlp(y)
        (y * w.to(dtype)).sum().backward()
        grads = {"g_angles": a_.grad}
        if fixed is None:
            grads["g_item_rep"] = it.grad
        if mlp is not None:
            grads.update({"g_mlp." + k: p.grad for k, p in mlp.named_parameters()})
        return y.detach(), grads
    y64, g64 = oracle(torch.float64)
    y32, g32 = oracle(torch.float32)
    as_good_as_ref32(out, y64, y32, "out")
    ours = {"g_angles": a.grad}
    if fixed is None:
        ours["g_item_rep"] = net.item_rep.grad
    if with_mlp:
        ours.update({"g_mlp." + k: p.grad for k, p in net.mlp.named_parameters()})
    assert set(ours) == set(g64)
    for k in g64:
        as_good_as_ref32(ours[k], g64[k], g32[k], k, floor=4.0 if k.startswith("g_mlp") else 1.0)   # cuBLAS (TF32-free) sums vs CPU order
