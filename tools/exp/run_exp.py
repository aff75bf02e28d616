#!/usr/bin/env python
"""Build and time the experimental Wigner variants (tools/exp/wigner_exp.cu)."""
import ctypes
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
SO = os.path.join(HERE, "libwigner_exp.so")


def build():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-shared", "-Xptxas", "-v", "-o", SO, os.path.join(HERE, "wigner_exp.cu")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode:
        print(p.stderr[-6000:])
        raise SystemExit(1)
    return p.stderr


if __name__ == "__main__":
    if "--build" in sys.argv:
        log = build()
        import re
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?Used (\d+) registers.*?\n", log, re.S):
            print(m.group(1)[:60], m.group(2))
        for ln in log.splitlines():
            if "spill" in ln and "0 bytes spill stores" not in ln:
                print(ln)
        raise SystemExit(0)
    import lie_vae_b200.lie_tools as lt
    lib = ctypes.CDLL(SO)
    B, L, C, M = 1 << 18, 8, 10, 81
    dev = torch.device("cuda")
    torch.manual_seed(0)
    ang = lt.group_matrix_to_eazyz(lt.random_group_matrices(B, device=dev))
    item = torch.randn(M, C, device=dev)
    ys = [torch.empty(B, M * C, device=dev) for _ in range(3)]
    ref = lt.block_wigner_matrix_multiply(ang, item.expand(B, -1, -1), L).view(B, -1)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    if sys.argv[1] == "bwdq":
        gy = [torch.randn(B, M * C, device=dev) for _ in range(3)]
        gang = torch.empty(B, 3, device=dev)
        gacc = torch.zeros(148 * 3 * 16 * M * C, device=dev)
        import lie_vae_b200._ops as ops
        a_ref = ang.clone().requires_grad_(True)
        it_ref = item.clone().requires_grad_(True)
        (ops.WignerApply.apply(a_ref, it_ref, 0, L, False).view(B, -1) * gy[0]).sum().backward()
        def run(i):
            rc = lib.exp_wigner_bwdq(P(ang), P(item), P(gy[i % 3]), P(gang), P(gacc), ctypes.c_int64(B), 148, st)
            assert rc == 0, rc
        run(0); torch.cuda.synchronize()
        gi = gacc.view(-1, M, C).sum(0)
        run(0); torch.cuda.synchronize()
        gi2 = gacc.view(-1, M, C).sum(0)
        err_a = float((gang - a_ref.grad).abs().max())
        err_i = float((gi - it_ref.grad).abs().max() / it_ref.grad.abs().max())
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            run(i)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print("bwdq: %.4f ms (incl. memset)  %.0f GB/s  err_angles %.2e err_item(rel) %.2e repeat_equal %s" % (ms, 3264 * B / ms / 1e6, err_a, err_i, bool(torch.equal(gi, gi2))), flush=True)
        raise SystemExit(0)
    if sys.argv[1] in ("bwdcm", "bwd2cm", "bwdws"):
        gy = [torch.randn(B, M * C, device=dev) for _ in range(3)]
        gang = torch.empty(B, 3, device=dev)
        part = torch.zeros(148 * 8 * M * C, device=dev)
        import lie_vae_b200._ops as ops
        a_ref = ang.clone().requires_grad_(True)
        it_ref = item.clone().requires_grad_(True)
        (ops.WignerApply.apply(a_ref, it_ref, 0, L, False).view(B, -1) * gy[0]).sum().backward()
        for gm in [int(v) for v in sys.argv[2].split(",")]:
            def run(i):
                fn = {"bwdcm": lib.exp_wigner_bwdcm, "bwd2cm": lib.exp_wigner_bwd2cm, "bwdws": lib.exp_wigner_bwdws}[sys.argv[1]]
                rc = fn(P(ang), P(item), P(gy[i % 3]), P(gang), P(part), ctypes.c_int64(B), 148 * gm, st)
                assert rc == 0, rc
            run(0)
            torch.cuda.synchronize()
            gi = part[:148 * gm * M * C].view(148 * gm, M, C).sum(0)
            err_a = float((gang - a_ref.grad).abs().max())
            err_i = float((gi - it_ref.grad).abs().max() / it_ref.grad.abs().max())
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(20):
                run(i)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print(sys.argv[1] + " grid 148x%d: %.4f ms  %.0f GB/s  err_angles %.2e err_item(rel) %.2e" % (gm, ms, 3264 * B / ms / 1e6, err_a, err_i), flush=True)
        raise SystemExit(0)
    if sys.argv[1] == "bwd2":
        gy = [torch.randn(B, M * C, device=dev) for _ in range(3)]
        gang = torch.empty(B, 3, device=dev)
        part = torch.zeros(148 * 8 * M * C, device=dev)
        import lie_vae_b200._ops as ops
        a_ref = ang.clone().requires_grad_(True)
        it_ref = item.clone().requires_grad_(True)
        (ops.WignerApply.apply(a_ref, it_ref, 0, L, False).view(B, -1) * gy[0]).sum().backward()
        for var in [int(v) for v in sys.argv[2].split(",")]:
            for S in [int(v) for v in sys.argv[3].split(",")]:
                for gm in [int(v) for v in sys.argv[4].split(",")]:
                    def run(i):
                        rc = lib.exp_wigner_bwd2(var, P(ang), P(item), P(gy[i % 3]), P(gang), P(part), ctypes.c_int64(B), S, 148 * gm, st)
                        assert rc == 0, rc
                    run(0)
                    torch.cuda.synchronize()
                    gi = part[:148 * gm * M * C].view(148 * gm, M, C).sum(0)
                    err_a = float((gang - a_ref.grad).abs().max())
                    err_i = float((gi - it_ref.grad).abs().max() / it_ref.grad.abs().max())
                    for i in range(3):
                        run(i)
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(20):
                        run(i)
                    b.record()
                    torch.cuda.synchronize()
                    ms = a.elapsed_time(b) / 20
                    print("bwd2 var %d S %d grid 148x%d: %.4f ms  %.0f GB/s  err_angles %.2e err_item(rel) %.2e" % (var, S, gm, ms, 3264 * B / ms / 1e6, err_a, err_i), flush=True)
        raise SystemExit(0)
    if sys.argv[1] == "bwdwarp":
        gy = [torch.randn(B, M * C, device=dev) for _ in range(3)]
        gang = torch.empty(B, 3, device=dev)
        part = torch.empty(148 * 8 * M * C, device=dev)
        # reference gradients from the product kernels
        import lie_vae_b200._ops as ops
        a_ref = ang.clone().requires_grad_(True)
        it_ref = item.clone().requires_grad_(True)
        (ops.WignerApply.apply(a_ref, it_ref, 0, L, False).view(B, -1) * gy[0]).sum().backward()
        for gm in [int(v) for v in sys.argv[2].split(",")]:
            for threads in [int(v) for v in sys.argv[3].split(",")]:
                def run(i):
                    rc = lib.exp_wigner_bwd_warp(0, P(ang), P(item), P(gy[i % 3]), P(gang), P(part), ctypes.c_int64(B), 148 * gm, threads, st)
                    assert rc == 0, rc
                run(0)
                torch.cuda.synchronize()
                gi = part[:148 * gm * M * C].view(148 * gm, M, C).sum(0)
                err_a = float((gang - a_ref.grad).abs().max())
                err_i = float((gi - it_ref.grad).abs().max() / it_ref.grad.abs().max())
                for i in range(3):
                    run(i)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(20):
                    run(i)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 20
                print("bwdwarp grid 148x%d threads %d: %.4f ms  %.0f GB/s  err_angles %.2e err_item(rel) %.2e" % (gm, threads, ms, 3264 * B / ms / 1e6, err_a, err_i), flush=True)
        raise SystemExit(0)
    if sys.argv[1] == "bwd":
        gy = [torch.randn(B, M * C, device=dev) for _ in range(3)]
        gang = torch.empty(B, 3, device=dev)
        part = torch.zeros(148 * 4 * 16 * M * C, device=dev)
        import lie_vae_b200._ops as ops
        gy0 = gy[0]
        a_ref = ang.clone().requires_grad_(True)
        it_ref = item.clone().requires_grad_(True)
        (ops.WignerApply.apply(a_ref, it_ref, 0, L, False).view(B, -1) * gy0).sum().backward()
        for var in [int(v) for v in sys.argv[2].split(",")]:
            for S in [int(v) for v in sys.argv[3].split(",")]:
                for gm in [int(v) for v in sys.argv[4].split(",")]:
                    def run(i):
                        rc = lib.exp_wigner_bwd(var, P(ang), P(item), P(gy[i % 3]), P(gang), P(part), ctypes.c_int64(B), S, 148 * gm, st)
                        assert rc == 0, rc
                    for i in range(3):
                        run(i)
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(20):
                        run(i)
                    b.record()
                    torch.cuda.synchronize()
                    ms = a.elapsed_time(b) / 20
                    extra = ""
                    if var & 128:
                        part.zero_(); run(0); torch.cuda.synchronize()
                        gi = part[:148 * gm * S * M * C].view(-1, M, C).sum(0)
                        gi2 = None
                        part.zero_(); run(0); torch.cuda.synchronize()
                        gi2 = part[:148 * gm * S * M * C].view(-1, M, C).sum(0)
                        extra = " err_item(rel) %.2e repeat_equal %s" % (float((gi - it_ref.grad).abs().max() / it_ref.grad.abs().max()), bool(torch.equal(gi, gi2)))
                    print("bwd var %3d S %2d grid 148x%d: %.4f ms  %.0f GB/s  chk %.4f%s" % (var, S, gm, ms, 3264 * B / ms / 1e6, float(gang.abs().mean()), extra), flush=True)
        raise SystemExit(0)
    for var in [int(v) for v in sys.argv[1].split(",")]:
        for S in [int(v) for v in sys.argv[2].split(",")]:
            def run(i):
                rc = lib.exp_wigner_fwd(var, P(ang), P(item), P(ys[i % 3]), ctypes.c_int64(B), S, st)
                assert rc == 0, rc
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(20):
                run(i)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            err = float((ys[0] - ref).abs().max()) if not (var & (1 | 8 | 16)) else float("nan")
            print("var %2d S %2d: %.4f ms  %.0f GB/s  err %.2e" % (var, S, ms, 3252 * B / ms / 1e6, err), flush=True)
