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


def test_normal_contract_numpy_equals_c_bit_for_bit():
    """The float32 Box-Muller of the RNG contract (oracle/philox.py normal_pairs32): the numpy restatement (with its
    emulated fused multiply-add) and the C oracle agree on every bit, including the corner words; known answers."""
    import ctypes as C
    from oracle import c_oracle
    lib = c_oracle.lib()
    rng = np.random.default_rng(7)
    n = 400000
    w0 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    w1 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    w0[:8] = [0, 255, 256, 2 ** 32 - 1, 2 ** 32 - 256, 2 ** 31, 12345, 2 ** 24]
    w1[:8] = [0, 2 ** 32 - 1, 2 ** 30, 2 ** 31, 3 * 2 ** 30, 2 ** 29, 2 ** 29 - 1, 2 ** 29 + 256]
    n0, n1 = px.normal_pairs32(w0, w1)
    c0, c1 = np.zeros(n, np.float32), np.zeros(n, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.dreamzs_oracle_normal_pairs32(p(w0), p(w1), C.c_int64(n), p(c0), p(c1))
    assert np.array_equal(n0.view(np.uint32), c0.view(np.uint32)) and np.array_equal(n1.view(np.uint32), c1.view(np.uint32))
    # known answers: u = 2^-24 (largest radius) at angle 0; u = 1 (radius 0); a quarter turn
    assert n0[0] == np.float32(5.7681074) and n1[0] == 0.0
    assert n0[3] == 0.0 and n1[3] == 0.0
    assert abs(float(n0[2])) < 1e-6 and abs(float(n1[2]) - 5.6466599) < 1e-6
    # it is a standard normal to float32 accuracy
    u = ((w0 >> np.uint32(8)).astype(np.float64) + 1) * 2.0 ** -24
    v = (w1 >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    r = np.sqrt(-2 * np.log(u))
    assert np.abs(n0 - r * np.cos(2 * np.pi * v)).max() < 2e-6 and np.abs(n1 - r * np.sin(2 * np.pi * v)).max() < 2e-6
    # the emulated fma against C's fmaf
    a, b = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    c = (rng.standard_normal(n) * 10.0 ** rng.integers(-8, 3, n)).astype(np.float32)
    out = np.zeros(n, np.float32)
    lib.dreamzs_oracle_fmaf(p(a), p(b), p(c), C.c_int64(n), p(out))
    assert np.array_equal(px.fma32(a, b, c).view(np.uint32), out.view(np.uint32))
