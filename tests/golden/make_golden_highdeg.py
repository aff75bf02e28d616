"""Generate the degree > 8 fixtures from the UNMODIFIED reference (float64); companion of make_golden.py.

    python tests/golden/make_golden_highdeg.py        (build container only: needs /root/reference)

block_wigner_L11C2_{N,T}.npz : lie_tools.block_wigner_matrix_multiply, degrees 0..11, per-sample spectrum
action_net_L10C4.npz        : decoders.ActionNet(degrees=10, rep_copies=4), shared item_rep
wigner_d_high.npz           : lie_tools.wigner_d_matrix for l = 9, 12, 16
They pin the oracle (and through it the generic sm_100a kernels) above the degrees the unrolled kernels cover.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (imports the reference through tests/refshim.py, float64 J)

lt, dc, F64 = mg.lt, mg.dc, mg.F64


def main():
    gen = torch.Generator().manual_seed(20181019)

    def randn(*s):
        return torch.randn(*s, generator=gen, dtype=F64)

    ang = lt.group_matrix_to_eazyz(lt.quaternions_to_group_matrix(randn(5, 4)))
    ang = torch.cat([ang, torch.tensor([[0.3, 1.1, -2.0]])], 0)
    mg.save("wigner_d_high", angles=ang, **{"D%d" % l: lt.wigner_d_matrix(ang, l) for l in (9, 12, 16)})
    L, C = 11, 2
    spec = randn(ang.shape[0], (L + 1) ** 2, C)
    for tr in (False, True):
        out, w, (ga, gs) = mg.grads(lambda a, s: lt.block_wigner_matrix_multiply(a, s, L, transpose=tr), [ang, spec], 21)
        mg.save("block_wigner_L11C2_%s" % ("T" if tr else "N"), angles=ang, spectrum=spec, out=out, w=w, gangles=ga,
                gspectrum=gs, max_degree=L)
    L, C = 10, 4
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C)
    item = randn((L + 1) ** 2, C)
    net.item_rep.data = item.clone()
    a = ang.clone().requires_grad_(True)
    out = net(a)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(22), dtype=F64)
    (out * w).sum().backward()
    mg.save("action_net_L10C4", angles=ang, item_rep=item, out=out, w=w, gangles=a.grad, gitem=net.item_rep.grad,
            degrees=L, transpose=0)


if __name__ == "__main__":
    main()
