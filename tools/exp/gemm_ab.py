import os, sys, json, torch
sys.path.insert(0, "/root/repo")
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
K = 810
for M, N in ((1024, 3200), (8192, 3200), (65536, 3200), (65536, 800)):
    a = torch.randn(M, K, device=dev)
    bt = _ops.round_tf32(torch.randn(N, 812, device=dev))[:, :K]
    out = torch.empty(M, N, device=dev)
    ms = timed(lambda: _ops.gemm_tf32(a, bt, out=out))
    ref = (a[:64].double() @ bt.double().t())
    err = float((out[:64].double() - ref).abs().max() / ref.pow(2).mean().sqrt())
    print(os.environ.get("LV_TAG", ""), M, N, round(ms, 4), "ms", round(2.0 * M * N * K / ms / 1e9, 1), "TFLOP/s  err", round(err, 5), flush=True)
