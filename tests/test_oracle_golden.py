"""The oracle (oracle/so3_oracle.py) against the reference's own outputs.

Fixtures in tests/golden/ were produced by the unmodified reference in float64
(tests/golden/make_golden.py).  The oracle must reproduce values AND autograd
gradients to 1e-12; this is what pins it.  CPU only.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import so3_oracle as O

TOL = dict(rtol=1e-11, atol=1e-12)


def T(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


def check_grads(fn, inputs, w, out_ref, grad_refs, tol=TOL):
    leaves = [T(x).requires_grad_(True) for x in inputs]
    out = fn(*leaves)
    np.testing.assert_allclose(out.detach().numpy(), out_ref, **tol)
    (out * T(w)).sum().backward()
    for leaf, g in zip(leaves, grad_refs):
        np.testing.assert_allclose(leaf.grad.numpy(), g, **tol)


def test_algebra():
    g = load_golden("algebra")
    np.testing.assert_allclose(O.map_to_lie_algebra(T(g["v"])).numpy(), g["hat"], **TOL)
    np.testing.assert_allclose(O.map_to_lie_vector(T(g["hat"])).numpy(), g["vee"], **TOL)


def test_rodrigues():
    g = load_golden("rodrigues")
    check_grads(O.rodrigues, [g["v"]], g["w"], g["out"], [g["gv"]], dict(rtol=1e-9, atol=1e-10))


def test_log_map():
    g = load_golden("log_map")
    check_grads(O.log_map, [g["R"]], g["w"], g["out"], [g["gR"]], dict(rtol=1e-10, atol=1e-11))


def test_quat_to_mat():
    g = load_golden("quat_to_mat")
    check_grads(O.quaternions_to_group_matrix, [g["q"]], g["w"], g["out"], [g["gq"]])


def test_mat_to_quat():
    g = load_golden("mat_to_quat")
    check_grads(O.group_matrix_to_quaternions, [g["R"]], g["w"], g["out"], [g["gR"]])


def test_quat_to_eazyz():
    g = load_golden("quat_to_eazyz")
    check_grads(O.quaternions_to_eazyz, [g["q"]], g["w"], g["out"], [g["gq"]])


def test_mat_to_eazyz():
    g = load_golden("mat_to_eazyz")
    check_grads(O.group_matrix_to_eazyz, [g["R"]], g["w"], g["out"], [g["gR"]], dict(rtol=1e-9, atol=1e-10))


def test_s2s1():
    g = load_golden("s2s1")
    check_grads(O.s2s1rodrigues, [g["s2"], g["s1"]], g["w"], g["out"], [g["gs2"], g["gs1"]])


def test_s2s2():
    g = load_golden("s2s2")
    check_grads(O.s2s2_gram_schmidt, [g["v1"], g["v2"]], g["w"], g["out"], [g["gv1"], g["gv2"]])


def test_vector_to_eazyz():
    g = load_golden("vector_to_eazyz")
    check_grads(O.vector_to_eazyz, [g["v"]], g["w"], g["out"], [g["gv"]])


def test_wigner_d():
    g = load_golden("wigner_d")
    for l in range(9):
        np.testing.assert_allclose(O.wigner_d_matrix(T(g["angles"]), l).numpy(), g["D%d" % l], **TOL)


def test_wigner_d_high_degree():
    g = load_golden("wigner_d_high")
    for l in (9, 12, 16):
        np.testing.assert_allclose(O.wigner_d_matrix(T(g["angles"]), l).numpy(), g["D%d" % l], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize("tag", ["L8C3", "L3C1", "L5C10", "L11C2"])
@pytest.mark.parametrize("tr", ["N", "T"])
def test_block_wigner(tag, tr):
    g = load_golden("block_wigner_%s_%s" % (tag, tr))
    L = int(g["max_degree"])
    check_grads(lambda a, s: O.block_wigner_matrix_multiply(a, s, L, transpose=(tr == "T")),
                [g["angles"], g["spectrum"]], g["w"], g["out"], [g["gangles"], g["gspectrum"]],
                dict(rtol=1e-10, atol=1e-11))


@pytest.mark.parametrize("name", ["action_net_L8C10", "action_net_L3C3", "action_net_L10C4"])
def test_action_net(name):
    g = load_golden(name)
    L, tr = int(g["degrees"]), bool(int(g["transpose"]))
    check_grads(lambda a, it: O.action_net_forward(a, it, L, tr), [g["angles"], g["item_rep"]], g["w"], g["out"],
                [g["gangles"], g["gitem"]], dict(rtol=1e-10, atol=1e-11))


@pytest.mark.parametrize("name", ["so3_reparam_k3", "so3_reparam_k10", "so3_reparam_n5", "so3_reparam_iwae",
                                  "so3_reparam_ka6"])
def test_so3_reparam(name):
    g = load_golden(name)
    k = int(g["k"])
    mu, sg, eps = T(g["mu"]).requires_grad_(True), T(g["sigma"]).requires_grad_(True), T(g["eps"])
    z, lq = O.so3_reparameterize(mu, sg, eps, k)
    np.testing.assert_allclose(z.detach().numpy(), g["z"], **TOL)
    np.testing.assert_allclose(lq.detach().numpy(), g["log_q"], rtol=1e-10, atol=1e-11)
    ((z * T(g["wz"])).sum() + (lq * T(g["wl"])).sum()).backward()
    np.testing.assert_allclose(mu.grad.numpy(), g["gmu"], **TOL)
    np.testing.assert_allclose(sg.grad.numpy(), g["gsigma"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(O.so3_log_prior(z).numpy(), g["log_prior"], **TOL)
    np.testing.assert_allclose(O.so3_kl(lq).detach().numpy(), g["kl"], rtol=1e-10, atol=1e-11)


def test_logsumexp():
    g = load_golden("logsumexp")
    check_grads(lambda t: O.logsumexp(t, 1), [g["x"]], g["w_dim1"], g["out_dim1"], [g["gx_dim1"]])
    check_grads(lambda t: O.logsumexp(t, 0), [g["x"]], g["w_dim0"], g["out_dim0"], [g["gx_dim0"]])
    np.testing.assert_allclose(O.logsumexp(T(g["x"])).numpy(), g["out_all"], **TOL)


# ---- known answers quoted in SURVEY.md Appendix B (independent of the npz files) ------------
def test_known_answers():
    R = O.rodrigues(T([0.1, 0.2, 0.3]))
    np.testing.assert_allclose(R.numpy(), [[0.9357548033, -0.2831649606, 0.2101917060],
                                           [0.3029327134, 0.9505806179, -0.0680313164],
                                           [-0.1805400767, 0.1273345749, 0.9752903090]], atol=1e-9)
    np.testing.assert_allclose(O.map_to_lie_vector(O.log_map(R)).numpy(), [0.1, 0.2, 0.3], atol=1e-12)
    M = O.quaternions_to_group_matrix(T([0.1, 0.2, 0.3, 0.4]))
    np.testing.assert_allclose(M.numpy(), [[2 / 15, 14 / 15, -1 / 3], [-2 / 3, 1 / 3, 2 / 3], [11 / 15, 2 / 15, 2 / 3]],
                               atol=1e-12)
    q = O.group_matrix_to_quaternions(M)
    np.testing.assert_allclose(q.numpy(), [0.1825741430, 0.3651482861, 0.5477224291, 0.7302969145], atol=1e-9)
    np.testing.assert_allclose(O.quaternions_to_eazyz(q).numpy(), [0.1798532748, 0.8410684190, 1.1071484928], atol=1e-9)
    np.testing.assert_allclose(O.group_matrix_to_eazyz(R).numpy(), [2.5273197923, 0.2227639061, -2.8285702969], atol=1e-9)
    g = load_golden("so3_reparam_ka6")
    np.testing.assert_allclose(g["log_q"][0], [-1.3132567320, -3.4611382701, 6.2003812210], atol=1e-9)
    ang = T([[0.3, 1.1, -2.0]])
    D2 = O.wigner_d_matrix(ang, 2)[0].numpy()
    np.testing.assert_allclose(D2[0], [-0.50232601, -0.09854251, 0.38838434, -0.76381725, 0.06081747], atol=2e-7)
    spec = T(np.arange(1, 10) / 10.0).reshape(1, 9, 1)
    np.testing.assert_allclose(O.block_wigner_matrix_multiply(ang, spec, 2).reshape(-1).numpy(),
                               [0.1, -0.3459106982, 0.1223535836, 0.3941769004, -0.5947375894, -0.5017371178,
                                -0.9477864504, 0.9462906122, 0.3883070350], atol=5e-7)
    np.testing.assert_allclose(O.block_wigner_matrix_multiply(ang, spec, 2, transpose=True).reshape(-1).numpy(),
                               [0.1, -0.0914413631, 0.5293139815, -0.0382780991, -0.3195415735, -0.6008644104,
                                1.2304338217, 0.2005974054, -0.7298277617], atol=5e-7)
