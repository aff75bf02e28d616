#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs[1] (isolated SO3 reparameterize fwd+bwd at 2^20 samples) and configs[2]
(Wigner-D action decoder, l <= 8, batch 65536, fwd+bwd) with their SURVEY 8(d) variants, through the public autograd
API.  CUDA events, rotating buffers larger than L2, 5 warm-ups + 20 timed iterations.  One JSON line per row."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lie_vae_b200.lie_tools as lt  # noqa: E402
import lie_vae_b200.reparameterize as rp  # noqa: E402
import lie_vae_b200.decoders as dc  # noqa: E402
from lie_vae_b200 import _cabi  # noqa: E402
from lie_vae_b200._ops import _stream  # noqa: E402

PEAK = 6557.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = torch.device("cuda")


def timed(fn, iters=20, warm=16):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def row(name, ms, samples, bytes_per_sample):
    gbs = samples * bytes_per_sample / ms / 1e6
    print(json.dumps({"config": name, "ms_fwd_bwd": round(ms, 4), "samples": samples, "samples_per_s": round(samples / ms * 1e3),
                      "bytes_per_sample": bytes_per_sample, "gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 4)}), flush=True)


torch.manual_seed(0)
# bring the clocks up before the first timed row (a fresh process starts from an idle GPU)
_w = torch.randn(8192, 8192, device=dev)
_t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_t[0].record()
while True:
    for _ in range(10):
        _w @ _w
    _t[1].record()
    torch.cuda.synchronize()
    if _t[0].elapsed_time(_t[1]) > 1500:
        break
del _w
# ---- configs[1]: isolated reparameterize, B = 2^20, n = 1 (248 B/sample fwd+bwd), 8 rotating input sets (> L2)
B, NSET = 1 << 20, 8
sets = [(lt.random_group_matrices(B, device=dev).requires_grad_(True),
         torch.nn.functional.softplus(torch.randn(B, 3, device=dev)).requires_grad_(True),
         torch.randn(1, B, 3, device=dev), torch.randn(1, B, 3, 3, device=dev), torch.randn(1, B, device=dev)) for _ in range(NSET)]
for k in (3, 10):
    def step(i, k=k):
        mu, sg, eps, gz, glq = sets[i % NSET]
        mu.grad = sg.grad = None
        z, lq = rp.so3_reparameterize(mu, sg, eps, k)
        torch.autograd.backward([z, lq], [gz, glq])
    row("configs[1] so3_reparameterize fwd+bwd, B=2^20, k=%d" % k, timed(step), B, 248)
# the same public function captured by torch.cuda.make_graphed_callables (lie_vae_b200/graphed.py): forward and backward
# replay as one graph launch each -- no Function.apply / engine / ctypes cost per call
import lie_vae_b200.graphed as gr  # noqa: E402
for k in (3, 10):
    mu0, sg0, eps0, _, _ = sets[0]
    g_fn = gr.graphed(lambda m, s, e, k=k: rp.so3_reparameterize(m, s, e, k),
                      (mu0.detach().clone().requires_grad_(True), sg0.detach().clone().requires_grad_(True), eps0.clone()))

    def step_g(i, g_fn=g_fn):
        mu, sg, eps, gz, glq = sets[i % NSET]
        mu.grad = sg.grad = None
        z, lq = g_fn(mu, sg, eps)
        torch.autograd.backward([z, lq], [gz, glq])
    row("configs[1] so3_reparameterize fwd+bwd, B=2^20, k=%d, public API under make_graphed_callables" % k, timed(step_g, iters=50), B, 248)
    del g_fn
sg_stress = [(0.02 + 2.48 * torch.rand(B, 3, device=dev)).requires_grad_(True) for _ in range(NSET)]


def step_stress(i):
    mu, _, eps, gz, glq = sets[i % NSET]
    sg = sg_stress[i % NSET]
    mu.grad = sg.grad = None
    z, lq = rp.so3_reparameterize(mu, sg, eps, 3)
    torch.autograd.backward([z, lq], [gz, glq])


row("configs[1] stress set sigma~U(0.02,2.5), k=3", timed(step_stress), B, 248)
# the same two launches through the C ABI on pre-allocated buffers (no autograd tape, no allocation): the kernels themselves
zb, lqb, gmub, gsgb = torch.empty(1, B, 3, 3, device=dev), torch.empty(1, B, device=dev), torch.empty(1, B, 3, 3, device=dev), torch.empty(1, B, 3, device=dev)
p_ = _cabi.ptr
for k in (3, 10):
    def step_c(i, k=k):
        mu, sg, eps, gz, glq = sets[i % NSET]
        st = _stream()
        _cabi.call("lv_so3_reparam_fwd_f32", p_(mu), p_(sg), p_(eps), p_(zb), p_(lqb), 1, B, k, st)
        _cabi.call("lv_so3_reparam_bwd_f32", p_(mu), p_(sg), p_(eps), p_(gz), p_(glq), p_(gmub), p_(gsgb), 1, B, k, st)
    row("configs[1] C ABI (lv_so3_reparam_fwd/bwd_f32, resident buffers), B=2^20, k=%d" % k, timed(step_c, iters=50), B, 248)
del sets, sg_stress
torch.cuda.empty_cache()

# ---- SO3reparameterize module (AlgebraMean, Din = 10) from encoder features, heads fused into the kernel or not
for Bm in (1024, 1 << 20):
    xs = [torch.randn(Bm, 10, device=dev).requires_grad_(True) for _ in range(4)]
    gzs, glqs = torch.randn(1, Bm, 3, 3, device=dev), torch.randn(1, Bm, device=dev)
    mod = rp.SO3reparameterize(rp.N0reparameterize(10, 3), rp.AlgebraMean(10), k=10).to(dev)
    for fuse in (True, False):
        mod.fuse_heads = fuse

        def step_m(i, mod=mod, xs=xs):
            x = xs[i % 4]
            x.grad = None
            mod.zero_grad(set_to_none=True)
            z = mod(x)
            torch.autograd.backward([z, mod.log_posterior()], [gzs, glqs])
        row("8f-2 SO3reparameterize(AlgebraMean) module fwd+bwd from features, B=%d, k=10, heads %s" % (Bm, "fused" if fuse else "separate launches"),
            timed(step_m), Bm, 40 + 12 + 36 + 4 + 36 + 4 + 40)
    del xs, gzs, glqs

# ---- configs[2]: ActionNet(degrees=8, rep_copies=10), N = 65536 (6516 B/sample shared, 16236 per-sample spectrum)
N, L = 65536, 8
M = (L + 1) ** 2
angs = [lt.group_matrix_to_eazyz(lt.random_group_matrices(N, device=dev)).requires_grad_(True) for _ in range(2)]
for C, tr in ((10, False), (10, True), (1, False)):
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C, transpose=tr).to(dev)
    gs = [torch.randn(N, M * C, device=dev) for _ in range(3)]

    def step(i, net=net, gs=gs):
        a = angs[i % 2]
        a.grad = None
        net.item_rep.grad = None
        net(a).backward(gs[i % 3])
    row("configs[2] ActionNet fwd+bwd, N=65536, l<=8, C=%d%s" % (C, ", transpose" if tr else ""), timed(step), N, 2 * 4 * M * C + 36)
    if C == 10 and not tr:
        g_net = gr.graphed(net, (angs[0].detach().clone().requires_grad_(True),))

        def step_gn(i, g_net=g_net, net=net, gs=gs):
            a = angs[i % 2]
            a.grad = None
            net.item_rep.grad = None
            g_net(a).backward(gs[i % 3])
        row("configs[2] ActionNet fwd+bwd, N=65536, l<=8, C=10, public API under make_graphed_callables", timed(step_gn, iters=50), N, 2 * 4 * M * C + 36)
        del g_net
    if C == 10:
        yb, gang, gitem = torch.empty(N, M * C, device=dev), torch.empty(N, 3, device=dev), torch.empty(M, C, device=dev)
        nws = _cabi.lib().lv_wigner_bwd_workspace_floats(N, 0, L, C)
        wsb = torch.empty(nws, device=dev)
        item = net.item_rep.detach()

        def step_c(i, tr=tr, gs=gs):
            st, a = _stream(), angs[i % 2].detach()
            _cabi.call("lv_wigner_apply_fwd_f32", p_(a), p_(item), p_(yb), N, 0, L, C, 1, int(tr), st)
            _cabi.call("lv_wigner_apply_bwd_f32", p_(a), p_(item), p_(gs[i % 3]), p_(gang), p_(gitem), p_(wsb), nws, N, 0, L, C, 1, int(tr), st)
        row("configs[2] C ABI (lv_wigner_apply_fwd/bwd_f32, resident buffers), C=10%s" % (", transpose" if tr else ""), timed(step_c, iters=50), N, 2 * 4 * M * C + 36)
C = 10
specs = [torch.randn(N, M, C, device=dev).requires_grad_(True) for _ in range(2)]
gs = [torch.randn(N, M, C, device=dev) for _ in range(2)]


def step_ps(i):
    a, s = angs[i % 2], specs[i % 2]
    a.grad = s.grad = None
    lt.block_wigner_matrix_multiply(a, s, L).backward(gs[i % 2])


row("configs[2] block_wigner_matrix_multiply fwd+bwd, per-sample spectrum (N,81,10)", timed(step_ps), N, 5 * 4 * M * C + 36)
del specs, gs, angs
torch.cuda.empty_cache()

# ---- the elementwise lie_tools maps through the C ABI at 2^22 rows (each tensor set > L2, two
# rotating sets); bytes = the map's inputs + outputs (forward), inputs + upstream gradient + input gradient (backward)
R = 1 << 22
mats = [lt.random_group_matrices(R, device=dev) for _ in range(2)]
ROWOPS = [("rodrigues", 3, 9, lambda: torch.randn(R, 3, device=dev)),
          ("log_map", 9, 9, None),
          ("quat_to_mat", 4, 9, lambda: torch.randn(R, 4, device=dev)),
          ("mat_to_quat", 9, 4, None),
          ("quat_to_eazyz", 4, 3, lambda: torch.randn(R, 4, device=dev)),
          ("mat_to_eazyz", 9, 3, None)]
for name, wi, wo, gen in ROWOPS:
    ins = mats if gen is None else [gen() for _ in range(2)]
    out, gout, gin = torch.empty(R, wo, device=dev), torch.randn(R, wo, device=dev), torch.empty(R, wi, device=dev)

    def f(i, name=name, ins=ins):
        _cabi.call("lv_%s_fwd_f32" % name, p_(ins[i % 2]), p_(out), R, _stream())

    def b(i, name=name, ins=ins):
        _cabi.call("lv_%s_bwd_f32" % name, p_(ins[i % 2]), p_(gout), p_(gin), R, _stream())
    row("8a lie_tools.%s forward, C ABI, 2^22 rows" % name, timed(f, iters=30), R, 4 * (wi + wo))
    row("8a lie_tools.%s backward, C ABI, 2^22 rows" % name, timed(b, iters=30), R, 4 * (2 * wi + wo))
    del ins, out, gout, gin
