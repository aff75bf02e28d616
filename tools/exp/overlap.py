#!/usr/bin/env python
"""Experiment: Wigner forward of micro-batch i+1 on a second stream, concurrent with the backward of micro-batch i
running on a reduced number of SMs (needs a build whose backward launch honours LV_EXP_BWD_CTAS; result, B200:
serial 0.373 ms per pair; overlapped with 132 / 116 / 100 backward CTAs 0.376 / 0.407 / 0.437 ms -- no gain: the
backward slows in proportion to its SMs and the concurrent forward stream delays its tile loads)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lie_vae_b200 import _cabi  # noqa: E402
import lie_vae_b200.lie_tools as lt  # noqa: E402

B, iters, L, C = 1 << 18, 24, 8, 10
M = (L + 1) ** 2
dev = torch.device("cuda")
torch.manual_seed(0)
ang = lt.group_matrix_to_eazyz(lt.random_group_matrices(B, device=dev))
item = torch.randn(M, C, device=dev)
gy = [torch.randn(B, M * C, device=dev) for _ in range(3)]
y = [torch.empty(B, M * C, device=dev) for _ in range(2)]
gang, gitem = torch.empty(B, 3, device=dev), torch.empty(M, C, device=dev)
nws = _cabi.lib().lv_wigner_bwd_workspace_floats(B, 0, L, C)
ws = torch.empty(nws, device=dev)
p = _cabi.ptr
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = lambda s: ctypes.c_void_p(s.cuda_stream)   # noqa: E731


def fwd(i, st):
    _cabi.call("lv_wigner_apply_fwd_f32", p(ang), p(item), p(y[i % 2]), B, 0, L, C, 1, 0, h(st))


def bwd(i, st):
    _cabi.call("lv_wigner_apply_bwd_f32", p(ang), p(item), p(gy[i % 3]), p(gang), p(gitem), p(ws), nws, B, 0, L, C, 1, 0, h(st))


def run(mode):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s1)
    for i in range(iters):
        if mode == "serial":
            fwd(i, s1)
            bwd(i, s1)
        else:                       # bwd first (takes its SMs), fwd fills the rest
            bwd(i, s1)
            fwd(i, s2)
    if mode != "serial":
        s1.wait_stream(s2)
    b.record(s1)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


mode = sys.argv[1] if len(sys.argv) > 1 else "serial"
run(mode)
print("ctas", os.environ.get("LV_EXP_BWD_CTAS", "all"), mode, "%.4f ms per (fwd + bwd) pair" % run(mode))
