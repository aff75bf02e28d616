"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference (float64).

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Every fixture stores the seeded inputs, the reference's outputs and, where the
function is differentiable, the reference's autograd gradients of the scalar
``sum(out * w)`` for stored random weights ``w``.  The reference is imported
through ``tests/refshim.py`` (stand-ins for three absent third-party packages;
the Wigner J table stand-in is the J-free construction of ``oracle/wigner_direct.py``; the committed files were
written when it still was the product's ``lie_vae_b200/jmatrix.py`` table -- regenerating with the independent one
reproduces them to 3.6e-13, they are held to 1e-11).
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refshim import load_reference  # noqa: E402

lt, rp, dc = load_reference()
lt._j_dtype = torch.float64
torch.set_default_dtype(torch.float64)
F64 = torch.float64


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (npy(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("wrote", path, {k: tuple(np.shape(v)) for k, v in arrs.items()})


def grads(fn, inputs, seed):
    """out = fn(*inputs); returns out, w, [d sum(out*w) / d input]."""
    g = torch.Generator().manual_seed(seed)
    leaves = [x.clone().requires_grad_(True) for x in inputs]
    out = fn(*leaves)
    w = torch.randn(out.shape, generator=g, dtype=F64)
    (out * w).sum().backward()
    return out.detach(), w, [x.grad.detach() for x in leaves]


def main():
    gen = torch.Generator().manual_seed(20181018)

    def randn(*s):
        return torch.randn(*s, generator=gen, dtype=F64)

    def rand(*s):
        return torch.rand(*s, generator=gen, dtype=F64)

    def rand_rot(n):
        q = randn(n, 4)
        return lt.quaternions_to_group_matrix(q)

    # ---- algebra maps -------------------------------------------------------------
    v = randn(16, 3)
    X = lt.map_to_lie_algebra(v)
    save("algebra", v=v, hat=X, vee=lt.map_to_lie_vector(X))

    # ---- rodrigues ----------------------------------------------------------------
    v = torch.cat([randn(24, 3) * 0.1, randn(24, 3), randn(16, 3) * 10.0,
                   torch.tensor([[0.1, 0.2, 0.3], [1e-4, -2e-4, 5e-5]])], 0)
    out, w, (gv,) = grads(lt.rodrigues, [v], 1)
    save("rodrigues", v=v, out=out, w=w, gv=gv)

    # ---- log_map (reference is single-matrix: loop) -------------------------------
    vs = torch.cat([randn(12, 3) * 0.3, randn(12, 3)], 0)
    vs = vs / vs.norm(dim=-1, keepdim=True).clamp(min=1.0) * vs.norm(dim=-1, keepdim=True).clamp(max=2.8)
    Rs = lt.rodrigues(vs)
    outs, ws, gRs = [], [], []
    for i in range(Rs.shape[0]):
        o, w, (gR,) = grads(lt.log_map, [Rs[i]], 100 + i)
        outs.append(o), ws.append(w), gRs.append(gR)
    save("log_map", R=Rs, out=torch.stack(outs), w=torch.stack(ws), gR=torch.stack(gRs))

    # ---- quaternion <-> matrix <-> Euler ------------------------------------------
    q = randn(64, 4) * (0.2 + 2 * rand(64, 1))
    out, w, (gq,) = grads(lt.quaternions_to_group_matrix, [q], 2)
    save("quat_to_mat", q=q, out=out, w=w, gq=gq)

    # all four Shepperd branches + perturbed (non-orthogonal) matrices for the gradient
    R = torch.cat([rand_rot(160),
                   lt.rodrigues(torch.tensor([[math.pi - 1e-3, 0, 0], [0, math.pi - 1e-3, 0],
                                              [0, 0, math.pi - 1e-3], [1e-3, 2e-3, -1e-3]])),
                   rand_rot(28) + 0.01 * randn(28, 3, 3)], 0)
    out, w, (gR,) = grads(lt.group_matrix_to_quaternions, [R], 3)
    branch = torch.stack([1 + R[:, 0, 0] - R[:, 1, 1] - R[:, 2, 2], 1 - R[:, 0, 0] + R[:, 1, 1] - R[:, 2, 2],
                          1 - R[:, 0, 0] - R[:, 1, 1] + R[:, 2, 2], 1 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]], 1).abs().argmax(1)
    assert set(branch.tolist()) == {0, 1, 2, 3}, "fixture must exercise every branch"
    save("mat_to_quat", R=R, out=out, w=w, gR=gR, branch=branch)

    qn = torch.cat([randn(60, 4), torch.tensor([[0.0, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1e-4],
                                                  [1e-4, 0.0, 1.0, 0.0], [0.5, 0.5, 0.5, 0.5]])], 0)
    qn = qn / qn.norm(dim=-1, keepdim=True)
    out, w, (gq,) = grads(lt.quaternions_to_eazyz, [qn], 4)
    save("quat_to_eazyz", q=qn, out=out, w=w, gq=gq)

    out, w, (gR,) = grads(lt.group_matrix_to_eazyz, [R], 5)
    save("mat_to_eazyz", R=R, out=out, w=w, gR=gR)

    # ---- S2xS1, S2xS2, tanh Euler --------------------------------------------------
    s2 = randn(32, 3)
    s2 = s2 / s2.norm(dim=-1, keepdim=True)
    s1 = randn(32, 2)
    s1 = s1 / s1.norm(dim=-1, keepdim=True)
    out, w, (g2, g1) = grads(lt.s2s1rodrigues, [s2, s1], 6)
    save("s2s1", s2=s2, s1=s1, out=out, w=w, gs2=g2, gs1=g1)

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        v1, v2 = randn(32, 3) * 10, randn(32, 3) * 10
        out, w, (g1, g2) = grads(lt.s2s2_gram_schmidt, [v1, v2], 7)
        save("s2s2", v1=v1, v2=v2, out=out, w=w, gv1=g1, gv2=g2)
        v = randn(32, 3) * 2
        out, w, (gv,) = grads(lt.vector_to_eazyz, [v], 8)
        save("vector_to_eazyz", v=v, out=out, w=w, gv=gv)

    # ---- Wigner --------------------------------------------------------------------
    ang = lt.group_matrix_to_eazyz(rand_rot(7))
    ang = torch.cat([ang, torch.tensor([[0.3, 1.1, -2.0]])], 0)
    wd = {"angles": ang}
    for l in range(9):
        wd["D%d" % l] = lt.wigner_d_matrix(ang, l)
    save("wigner_d", **wd)

    for tag, L, C in [("L8C3", 8, 3), ("L3C1", 3, 1), ("L5C10", 5, 10)]:
        M = (L + 1) ** 2
        spec = randn(ang.shape[0], M, C)
        for tr in (False, True):
            out, w, (ga, gs) = grads(lambda a, s: lt.block_wigner_matrix_multiply(a, s, L, transpose=tr), [ang, spec], 9)
            save("block_wigner_%s_%s" % (tag, "T" if tr else "N"), angles=ang, spectrum=spec, out=out, w=w, gangles=ga,
                 gspectrum=gs, max_degree=L)

    for L, C, tr in [(8, 10, False), (3, 3, True)]:
        net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, transpose=tr)
        item = randn((L + 1) ** 2, C)
        net.item_rep.data = item.clone()
        a = ang.clone().requires_grad_(True)
        out = net(a)
        g = torch.Generator().manual_seed(10)
        w = torch.randn(out.shape, generator=g, dtype=F64)
        (out * w).sum().backward()
        save("action_net_L%dC%d" % (L, C), angles=ang, item_rep=item, out=out, w=w, gangles=a.grad,
             gitem=net.item_rep.grad, degrees=L, transpose=int(tr))

    # ---- SO3 reparameterize + wrapped log-density ----------------------------------
    def so3_case(name, mu, sigma, eps, k):
        m = rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.AlgebraMean(10), k=k)
        mu_l = mu.clone().requires_grad_(True)
        sg_l = sigma.clone().requires_grad_(True)
        n = eps.shape[0]
        m.mu_lie = mu_l
        m.reparameterize.sigma = sg_l
        m.reparameterize.z = eps * sg_l
        m.v = m.reparameterize.z
        z = m.nsample(n)
        m.z = z
        lq = m.log_posterior()
        g = torch.Generator().manual_seed(11)
        wz = torch.randn(z.shape, generator=g, dtype=F64)
        wl = torch.randn(lq.shape, generator=g, dtype=F64)
        ((z * wz).sum() + (lq * wl).sum()).backward()
        save(name, mu=mu, sigma=sigma, eps=eps, k=k, z=z, log_q=lq, wz=wz, wl=wl, gmu=mu_l.grad, gsigma=sg_l.grad,
             log_prior=m.log_prior(), kl=m.kl())

    B = 96
    mu = rand_rot(B)
    sigma = torch.nn.functional.softplus(randn(B, 3))
    sigma[64:80] = 0.02 + 0.2 * rand(16, 3)            # sharp posteriors
    sigma[80:] = 1.0 + 1.5 * rand(16, 3)               # wide: winding terms matter
    eps = randn(1, B, 3)
    eps[0, 60:64] *= 1e-2                              # theta below sqrt(1e-3): both clamps active
    so3_case("so3_reparam_k3", mu, sigma, eps, 3)
    so3_case("so3_reparam_k10", mu, sigma, eps, 10)
    so3_case("so3_reparam_n5", mu[:8], sigma[:8], randn(5, 8, 3), 3)
    so3_case("so3_reparam_iwae", mu[:1], sigma[:1], randn(37, 1, 3), 10)

    # ---- known-answer vectors -------------------------------
    ka_sigma = torch.tensor([[.4, .5, .6], [1.5, 2, 2.5], [.05, .05, .05]])
    ka_eps = torch.tensor([[[.75, -.4, 5 / 6], [.2, -.1, .3], [.1, .2, -.1]]])
    ka_mu = lt.quaternions_to_group_matrix(torch.tensor([0.1, 0.2, 0.3, 0.4])).expand(3, 3, 3).contiguous()
    so3_case("so3_reparam_ka6", ka_mu, ka_sigma, ka_eps, 3)

    # ---- logsumexp -------------------------------------------------------------------
    from lie_vae.utils import logsumexp
    x = randn(5, 7, 3) * 4
    out, w, (gx,) = grads(lambda t: logsumexp(t, 1), [x], 12)
    out0, w0, (gx0,) = grads(lambda t: logsumexp(t, 0), [x], 13)
    save("logsumexp", x=x, out_dim1=out, w_dim1=w, gx_dim1=gx, out_dim0=out0, w_dim0=w0, gx_dim0=gx0,
         out_all=logsumexp(x))


if __name__ == "__main__":
    main()
