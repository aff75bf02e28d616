"""A/B of the tcgen05 GEMM's tile width on the shapes the fused action op launches (LV_GEMM_BN=128|256 overrides the choice)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lie_vae_b200 import _ops
dev = torch.device("cuda")
def timed(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
torch.manual_seed(0)
SHAPES = [(1024, 3200, 810, "fwd configs[3]-like L8"), (1024, 3200, 490, "fwd configs[3] L6"), (1024, 490, 3200, "dgrad configs[3] L6"), (1024, 810, 3200, "dgrad L8 hidden 200"),
          (8192, 3200, 810, "fwd chunk"), (8192, 810, 3200, "dgrad chunk hidden 200"), (8192, 800, 810, "fwd chunk hidden 50"), (8192, 810, 800, "dgrad chunk hidden 50"),
          (65536, 3200, 810, "large")]
for M, N, K, what in SHAPES:
    a = torch.randn(M, K + (K & 1), device=dev)[:, :K]
    Kp = (K + 3) // 4 * 4
    bt = _ops.round_tf32(torch.randn(N, Kp, device=dev))[:, :K]
    out = torch.empty(M, N, device=dev)
    res = []
    for bn in ("128", "256", ""):
        if bn: os.environ["LV_GEMM_BN"] = bn
        else: os.environ.pop("LV_GEMM_BN", None)
        ms = timed(lambda: _ops.gemm_tf32(a, bt, out=out))
        res.append("%s %.4f ms %.0f TF/s" % (bn or "auto", ms, 2.0 * M * N * K / ms / 1e9))
    ref = (a[:64].double() @ bt.double().t())
    err = float((out[:64].double() - ref).abs().max() / ref.pow(2).mean().sqrt())
    print("%-28s M=%d N=%d K=%d: %s  err %.5f" % (what, M, N, K, " | ".join(res), err), flush=True)
