"""Philox4x32-10 + Box-Muller in numpy: CPU restatement of the in-kernel noise of the fused reparameterize kernels
(``lie_vae_b200/csrc/reparam_core.cuh``: philox4x32_10 / philox_normal3).  TEST INFRASTRUCTURE (see so3_oracle.py).

The reference draws ``Normal(0,1).sample((n,))`` from torch's global generator (``lie_vae/reparameterize.py:137-141``) and
pins no stream; the in-kernel generator is this repo's own, so its oracle is the published Philox4x32-10 algorithm
(Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 known-answer vectors in
tests/test_philox_cpu.py) followed by the Box-Muller arithmetic the kernel states.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            h0, l0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            h1, l1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [h1 ^ c[1] ^ k0, l1, h0 ^ c[3] ^ k1, l0]
            k0 = k0 + W0
            k1 = k1 + W1
    return np.stack(c, axis=-1)


def philox_normal3(rows, seed, offset=0):
    """(rows, 3) float32: eps of flat samples offset .. offset + rows - 1 under key ``seed`` (counter = sample index)."""
    idx = np.arange(rows, dtype=np.uint64) + np.uint64(offset)
    ctr = np.zeros((rows, 4), dtype=np.uint32)
    ctr[:, 0] = (idx & MASK).astype(np.uint32)
    ctr[:, 1] = (idx >> np.uint64(32)).astype(np.uint32)
    seed = np.uint64(seed)
    key = np.broadcast_to(np.array([seed & MASK, seed >> np.uint64(32)], dtype=np.uint64).astype(np.uint32), (rows, 2))
    x = philox4x32_10(ctr, key)
    f = np.float32
    u0 = ((x[:, 0] >> np.uint32(8)).astype(f) + f(0.5)) * f(2.0 ** -24)
    u2 = ((x[:, 2] >> np.uint32(8)).astype(f) + f(0.5)) * f(2.0 ** -24)
    a1 = x[:, 1].astype(f) * f(2.0 ** -31)
    a3 = x[:, 3].astype(f) * f(2.0 ** -31)
    r0 = np.sqrt(f(-2.0) * np.log(u0), dtype=f)
    r2 = np.sqrt(f(-2.0) * np.log(u2), dtype=f)
    out = np.empty((rows, 3), dtype=f)
    out[:, 0] = r0 * np.cos(np.pi * a1.astype(np.float64)).astype(f)
    out[:, 1] = r0 * np.sin(np.pi * a1.astype(np.float64)).astype(f)
    out[:, 2] = r2 * np.cos(np.pi * a3.astype(np.float64)).astype(f)
    return out
