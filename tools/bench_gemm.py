#!/usr/bin/env python
"""The action's consumer: lv_gemm_tf32_f32 (tcgen05, TF32) against cuBLAS TF32 / FP32 on the same shapes,
and ActionNet + first DeconvNet layer, fused (L2-resident chunks) vs unfused, forward + backward.  One JSON line per row."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lie_vae_b200 import _ops  # noqa: E402
import lie_vae_b200.decoders as dc  # noqa: E402
import lie_vae_b200.lie_tools as lt  # noqa: E402

dev = torch.device("cuda")


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


torch.manual_seed(0)
K = 810
for M, N in ((1024, 3200), (8192, 3200), (65536, 3200), (65536, 800), (8192, 800)):
    a = torch.randn(M, K, device=dev)
    bt = _ops.round_tf32(torch.randn(N, 812, device=dev))[:, :K]
    b = bt.t().contiguous()
    out = torch.empty(M, N, device=dev)
    flops = 2.0 * M * N * K
    ms = timed(lambda: _ops.gemm_tf32(a, bt, out=out))
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32 = timed(lambda: torch.matmul(a, b, out=out))
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_fp32 = timed(lambda: torch.matmul(a, b, out=out))
    print(json.dumps({"row": "gemm M=%d N=%d K=%d" % (M, N, K), "lv_gemm_tf32_ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1),
                      "cublas_tf32_ms": round(ms_tf32, 4), "cublas_tf32_tflops": round(flops / ms_tf32 / 1e9, 1),
                      "cublas_fp32_ms": round(ms_fp32, 4)}), flush=True)


class View(torch.nn.Module):
    def __init__(self, *v):
        super().__init__()
        self.v = v

    def forward(self, x):
        return x.view(*self.v)


for L, hidden, N in ((6, 200, 1024), (8, 200, 65536), (8, 50, 65536), (8, 50, 1 << 18)):
    Mh = (L + 1) ** 2
    deconv = torch.nn.Sequential(View(-1, Mh * 10, 1, 1), torch.nn.ConvTranspose2d(Mh * 10, hidden, 4, 1, 0))
    net = dc.ActionNet(L, deconv, rep_copies=10).to(dev)
    ang = lt.group_matrix_to_eazyz(lt.random_group_matrices(N, device=dev)).requires_grad_(True)
    g = torch.randn(N, hidden, 4, 4, device=dev)
    res = {}
    for fuse in (False, True):
        net.fuse_consumer = fuse

        def step():
            ang.grad = None
            net.zero_grad(set_to_none=True)
            net(ang).backward(g)

        def fwd():
            with torch.no_grad():
                net(ang)
        res["fused" if fuse else "unfused"] = (round(timed(fwd, iters=10, warm=3), 4), round(timed(step, iters=10, warm=3), 4))
    print(json.dumps({"row": "ActionNet(l<=%d) + ConvTranspose2d(%d->%d,4,1,0), N=%d: ms (forward, forward+backward)" % (L, Mh * 10, hidden, N),
                      "unfused": res["unfused"], "fused_l2_chunks_tcgen05": res["fused"]}), flush=True)
