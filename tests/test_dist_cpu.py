"""world_size-2 gloo test of the multi-GPU host logic (CPU): sharding + the single all-reduce.

Each rank evaluates its batch shard with the ORACLE (test infrastructure; the CUDA kernels need a GPU),
packs [loss, grad item_rep], all-reduces, and the result must equal the single-process evaluation of the
concatenated batch (N-rank == 1-rank to 1e-6 relative, sum-order tolerance).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from lie_vae_b200.dist import shard_bounds, pack_reduction, unpack_reduction, allreduce_loss_and_grad  # noqa: E402


def test_shard_bounds_cover_and_align():
    for total, world, mult in [(1 << 24, 8, 1 << 18), (1000, 3, 1), (1000, 3, 64), (5, 8, 1), (0, 2, 4), (777, 1, 256)]:
        covered = 0
        prev_hi = 0
        for r in range(world):
            lo, hi = shard_bounds(total, world, r, mult)
            assert lo == prev_hi and hi >= lo
            if r < world - 1:
                assert (hi - lo) % mult == 0
            covered += hi - lo
            prev_hi = hi
        assert covered == total
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_pack_unpack_roundtrip():
    g = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    buf = pack_reduction(torch.tensor(2.5), g)
    assert buf.shape == (13,) and buf[0] == 2.5
    loss, gg = unpack_reduction(buf, g.shape)
    assert float(loss) == 2.5 and torch.equal(gg, g)
    with pytest.raises(ValueError):
        pack_reduction(torch.tensor(0.0), g, out=torch.empty(5))
    # without a process group the all-reduce is the identity
    loss, gg = allreduce_loss_and_grad(torch.tensor(1.0), g)
    assert float(loss) == 1.0 and torch.equal(gg, g)


def _local_step(lo, hi, L, C, k):
    from oracle import so3_oracle as O
    g = torch.Generator().manual_seed(7)
    B = 96
    mu = O.random_group_matrices(B, dtype=torch.float64, generator=g)
    sigma = torch.nn.functional.softplus(torch.randn(B, 3, dtype=torch.float64, generator=g))
    eps = torch.randn(1, B, 3, dtype=torch.float64, generator=g)
    item = torch.randn((L + 1) ** 2, C, dtype=torch.float64, generator=g).requires_grad_(True)
    gy = torch.randn(B, (L + 1) ** 2 * C, dtype=torch.float64, generator=g)
    glq = torch.randn(1, B, dtype=torch.float64, generator=g)
    z, lq = O.so3_reparameterize(mu[lo:hi], sigma[lo:hi], eps[:, lo:hi], k)
    y = O.action_net_forward(O.group_matrix_to_eazyz(z[0]), item, L)
    loss = (y * gy[lo:hi]).sum() + (lq * glq[:, lo:hi]).sum()
    loss.backward()
    return loss.detach().float(), item.grad.float()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(96, world, rank, 8)
        loss, gitem = _local_step(lo, hi, 3, 2, 3)
        loss, gitem = allreduce_loss_and_grad(loss, gitem)
        q.put((rank, float(loss), gitem.numpy().copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(180)
def test_two_rank_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    ref_loss, ref_g = _local_step(0, 96, 3, 2, 3)
    for _, loss, g in results:
        assert abs(loss - float(ref_loss)) <= 1e-5 * abs(float(ref_loss)) + 1e-5
        np.testing.assert_allclose(g, ref_g.numpy(), rtol=1e-5, atol=1e-5)
