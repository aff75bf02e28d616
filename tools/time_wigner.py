#!/usr/bin/env python
"""Time the Wigner forward / backward launches in isolation (CUDA events, rotating buffers > L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lie_vae_b200 import _cabi  # noqa: E402
from lie_vae_b200._ops import _stream  # noqa: E402
import lie_vae_b200.lie_tools as lt  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L, C = int(os.environ.get("LV_L", "8")), 10
M = (L + 1) ** 2
dev = torch.device("cuda")
torch.manual_seed(0)
ang = lt.group_matrix_to_eazyz(lt.random_group_matrices(B, device=dev))
item = torch.randn(M, C, device=dev)
NB = 3
gy = [torch.randn(B, M * C, device=dev) for _ in range(NB)]
y = [torch.empty(B, M * C, device=dev) for _ in range(NB)]
gang = torch.empty(B, 3, device=dev)
gitem = torch.empty(M, C, device=dev)
nws = _cabi.lib().lv_wigner_bwd_workspace_floats(B, 0, L, C)
ws = torch.empty(nws, device=dev)
p, st = _cabi.ptr, _stream()


def fwd(i):
    _cabi.call("lv_wigner_apply_fwd_f32", p(ang), p(item), p(y[i % NB]), B, 0, L, C, 1, 0, st)


def bwd(i):
    _cabi.call("lv_wigner_apply_bwd_f32", p(ang), p(item), p(gy[i % NB]), p(gang), p(gitem), p(ws), nws, B, 0, L, C, 1, 0, st)


for name, fn, nbytes in (("fwd", fwd, 3252), ("bwd", bwd, 3264)):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print("%s %s: %.4f ms  %.0f GB/s  (%.3f of 6557)  chk %.6f" % (os.environ.get("LV_TAG", ""), name, ms, nbytes * B / ms / 1e6,
                                                              nbytes * B / ms / 1e6 / 6557.4, float(y[0].abs().mean() if name == "fwd" else gitem.abs().mean())))
