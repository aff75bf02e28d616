"""N ranks compute what one rank computes -- on the GPU, through the kernels (last paragraph).

The sharded step of bench.py (contiguous batch slices, inputs generated per global chunk, upstream-gradient buffers indexed
by the global micro-batch number, one all-reduce of [loss, grad item_rep]) must give the single-GPU loss and item_rep
gradient to 1e-6 relative (summation order only):
  * always: the shards of world sizes 2 and 4 evaluated one after the other on this GPU and summed on the host;
  * with >= 2 GPUs: two real ranks over NCCL (``lie_vae_b200.dist.allreduce_packed``).
``-m gpu``.
"""
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

TOTAL, MICRO, L, C, K = 1 << 20, 1 << 18, 8, 10, 3


def shard_step(world, rank, dev):
    """[loss, grad item_rep] of rank ``rank`` of ``world`` for the TOTAL-sample step, as bench.py computes it."""
    import bench
    import lie_vae_b200.lie_tools as lt
    from lie_vae_b200 import dist as lvdist
    from lie_vae_b200.pipeline import FusedSO3ActionStep
    M = (L + 1) ** 2
    lo, hi = lvdist.shard_bounds(TOTAL, world, rank, MICRO)
    n_loc, n_micro, g_first = hi - lo, (hi - lo) // MICRO, lo // MICRO
    mu, sigma, eps, glq = bench.gen_inputs(lo, hi, dev, lt)
    item = torch.randn(M, C, device=dev, generator=torch.Generator(device=dev).manual_seed(0xA11CE))
    gy = [torch.randn(MICRO, M * C, device=dev, generator=torch.Generator(device=dev).manual_seed(0xD0 + j)) for j in range(3)]
    y = torch.empty(MICRO, M * C, device=dev)
    log_q, g_mu, g_sigma = torch.empty(n_loc, device=dev), torch.empty(n_loc, 3, 3, device=dev), torch.empty(n_loc, 3, device=dev)
    st = FusedSO3ActionStep(n_loc, MICRO, L, C, K, device=dev)
    st.g_item.zero_()
    st.latent_forward(mu, sigma, eps, log_q)
    for i in range(n_micro):
        st.decode_forward(i * MICRO, (i + 1) * MICRO, item, y)
        st.decode_backward(i * MICRO, (i + 1) * MICRO, item, gy[(g_first + i) % 3], accumulate=True)
    st.latent_backward(mu, sigma, eps, glq, g_mu, g_sigma)
    red = lvdist.pack_reduction(bench.step_loss(item, st.g_item, log_q, glq), st.g_item)
    return red, (g_mu, g_sigma, lo, hi)


def check_equal(red_n, red_1, what):
    loss_n, loss_1 = float(red_n[0]), float(red_1[0])
    assert abs(loss_n - loss_1) <= 1e-6 * max(1.0, abs(loss_1)) + 1e-6 * float(red_1[1:].abs().max()), (what, loss_n, loss_1)
    scale = float(red_1[1:].abs().max())
    assert float((red_n[1:] - red_1[1:]).abs().max()) <= 2e-6 * scale, what


def test_virtual_ranks_sum_to_single_rank():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda", 0)
    red_1, (gmu_1, gsg_1, _, _) = shard_step(1, 0, dev)
    for world in (2, 4):
        acc = torch.zeros_like(red_1, dtype=torch.float64)
        for r in range(world):
            red_r, (gmu_r, gsg_r, lo, hi) = shard_step(world, r, dev)
            acc += red_r.double()
            # per-sample gradients stay on the owning rank and are the single-rank ones, bit for bit
            assert torch.equal(gmu_r, gmu_1[lo:hi]) and torch.equal(gsg_r, gsg_1[lo:hi])
        check_equal(acc.float(), red_1, "world=%d" % world)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from lie_vae_b200 import dist as lvdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        red, _ = shard_step(world, rank, dev)
        lvdist.allreduce_packed(red)
        torch.cuda.synchronize()
        if rank == 0:
            red_1, _ = shard_step(1, 0, dev)
            q.put((red.cpu(), red_1.cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_nccl_ranks_match_single_rank():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    red_2, red_1 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    check_equal(red_2, red_1, "nccl world=2")
