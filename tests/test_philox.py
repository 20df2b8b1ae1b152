"""Known-answer tests of the RNG contract (Random123 kat_vectors for philox4x32-10) and agreement of the
numpy-vectorised, scalar and C implementations."""
import ctypes as C

import numpy as np

from oracle import philox as px


def test_philox_known_answers():
    assert px.philox4x32(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    m = 0xffffffff
    assert px.philox4x32(m, m, m, m, m, m) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert px.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_vectorised_matches_scalar():
    b = px.philox4x32_blocks(5, 11, 12, 13, 14, 15)
    for i in range(5):
        assert tuple(int(x) for x in b[i]) == px.philox4x32(i, 11, 12, 13, 14, 15)


def test_sample_distinct_and_uniform():
    s = px.Stream(3, 1, 2)
    counts = np.zeros(7)
    for it in range(4000):
        s = px.Stream(3, 1, it)
        r = s.sample(7, 4)
        assert len(set(r)) == 4 and all(0 <= x < 7 for x in r)
        counts[r] += 1
    assert np.all(np.abs(counts / counts.sum() - 1 / 7) < 0.01)


def test_normal_vec_moments():
    x = np.concatenate([px.Stream(9, c, 0).normal_vec(400) for c in range(100)])
    assert abs(x.mean()) < 0.02 and abs(x.std() - 1) < 0.02
