"""Experiment: how much of the power-capped slowdown of the Wigner backward is due to the forward's power draw?
Runs the bench's inner loop (fwd, bwd alternating over 16 micro-batches, 6 steps) with the real forward vs a pure-store stand-in."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from lie_vae_b200.pipeline import FusedSO3ActionStep
import lie_vae_b200.lie_tools as lt
dev = torch.device("cuda")
micro, nm, L, C = 1 << 20, 16, 8, 10
M = (L + 1) ** 2
step = FusedSO3ActionStep(micro * nm, micro, L, C, 3, device=dev)
step.angles.copy_(torch.rand(micro * nm, 3, device=dev) * 6 - 3)
item = torch.randn(M, C, device=dev)
gy = [torch.randn(micro, M * C, device=dev) for _ in range(3)]
y = [torch.empty(micro, M * C, device=dev) for _ in range(2)]
for mode in ("real", "store_only", "none", "real"):
    for rep in range(2):
        ev = []
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in range(6):
            for i in range(nm):
                a, b = i * micro, (i + 1) * micro
                if mode == "real":
                    step.decode_forward(a, b, item, y[i % 2])
                elif mode == "store_only":
                    y[i % 2].fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step.decode_backward(a, b, item, gy[i % 3], accumulate=True)
                e1.record()
                ev.append((e0, e1))
        t1.record()
        torch.cuda.synchronize()
        bw = [a.elapsed_time(b) for a, b in ev]
        print("%-10s total %.2f ms/step   bwd avg %.4f ms (first 16: %.4f, last 32: %.4f)" % (mode, t0.elapsed_time(t1) / 6, sum(bw) / len(bw), sum(bw[:16]) / 16, sum(bw[-32:]) / 32), flush=True)
