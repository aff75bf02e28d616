"""CPU tests of the Philox restatement (oracle/philox.py) that pins the in-kernel noise of the fused reparameterize kernels."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import philox as P  # noqa: E402


def _u32(*v):
    return np.array(v, dtype=np.uint32)


def test_philox4x32_10_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors)."""
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        got = P.philox4x32_10(_u32(*ctr), _u32(*key))
        assert tuple(int(g) for g in got) == want, (ctr, key, [hex(int(g)) for g in got])


def test_philox_normal_statistics_and_offsets():
    e = P.philox_normal3(1 << 16, seed=1234, offset=7)
    assert e.shape == (1 << 16, 3) and e.dtype == np.float32 and np.isfinite(e).all()
    assert abs(e.mean()) < 0.01 and abs(e.std() - 1.0) < 0.01
    assert abs(np.mean(e ** 3)) < 0.03 and abs(np.mean(e ** 4) - 3.0) < 0.1
    c = np.corrcoef(e.T)
    assert np.abs(c - np.eye(3)).max() < 0.02
    # the stream is indexed by the flat sample index: a shifted window is the same numbers
    assert np.array_equal(P.philox_normal3(100, 1234, 57), e[50:150])
    assert not np.array_equal(P.philox_normal3(100, 1235, 57), e[50:150])
