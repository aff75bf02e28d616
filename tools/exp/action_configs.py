import sys, torch
sys.path.insert(0, '/root/repo')
import lie_vae_b200.lie_tools as lt, lie_vae_b200.decoders as dc
dev = "cuda"
N = 65536
ang = [lt.group_matrix_to_eazyz(lt.random_group_matrices(N, device=dev)).requires_grad_(True) for _ in range(2)]
def timed(fn, iters=30, warm=10):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for L, C in ((8, 10), (6, 10), (4, 10), (8, 16), (8, 4), (3, 3), (6, 32)):
    M = (L + 1) ** 2
    net = dc.ActionNet(L, torch.nn.Sequential(), rep_copies=C).to(dev)
    gs = [torch.randn(N, M * C, device=dev) for _ in range(3)]
    def step(i):
        a = ang[i % 2]; a.grad = None; net.item_rep.grad = None
        net(a).backward(gs[i % 3])
    ms = timed(step)
    print("L=%d C=%d: %.4f ms fwd+bwd, %.0f GB/s algorithmic (%.0f B/sample)" % (L, C, ms, N * (8 * M * C + 36) / ms / 1e6, 8 * M * C + 36))
