"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Counter-based RNG spec shared by the oracle, the lock-step reference harness
and the CUDA kernels: Philox4x32-10 (Salmon et al., SC'11 "Parallel random
numbers: as easy as 1, 2, 3"; constants as published in Random123).

    key      = (seed & 0xffffffff, seed >> 32)
    counter  = (block, (call_no << 3) | stream, iteration, global_chain_id)

``stream`` identifies which RNG primitive of the reference's ``Dream.py`` is being
served and ``call_no`` is the running number of calls to that primitive inside
the current ``astep`` (per chain, reset every iteration).  ``block`` indexes
successive 4x32-bit outputs of one call.  The mapping primitive -> stream is

    0 MULTINOMIAL   np.random.multinomial(1, p)         Dream.py:545,565,595,615,908
    1 SAMPLE        random.sample(range(M), n)          Dream.py:662-664
    2 NORMAL        np.random.normal(0, s, d)           Dream.py:694
    3 UNIFORM_VEC   np.random.uniform(lo, hi, size)     Dream.py:696,700
    4 UNIFORM_SCAL  np.random.uniform([lo, hi])         Dream.py:618,993
    5 RAND          np.random.rand(m)                   Dream.py:749-751,773-775
    6 RANDINT       np.random.randint(1, n+1, size=1)   Dream.py:580

The parent process of the parallel-tempering driver (pydream/core.py:131-236) draws once per
iteration from the stream of the pseudo-chain ``DRIVER_CHAIN`` = 0xFFFFFFFF:
``np.random.choice(nchains, 2, replace=False)`` (core.py:183) = ``Stream.sample(nchains, 2)``
(stream 1, first pick then second pick) and ``np.random.uniform()`` (core.py:195) = stream 4.
"""
import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

ST_MULTINOMIAL, ST_SAMPLE, ST_NORMAL, ST_UNIFORM_VEC, ST_UNIFORM_SCAL, ST_RAND, ST_RANDINT = range(7)
DRIVER_CHAIN = 0xFFFFFFFF

TWO_M32 = 2.0 ** -32
TWO_M53 = 2.0 ** -53


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Scalar Philox4x32-10 on Python ints; returns 4 uint32 words."""
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_blocks(nblocks, c1, c2, c3, k0, k1):
    """Vectorised over block index c0 = 0..nblocks-1; returns uint64 array (nblocks, 4) of 32-bit words."""
    c0 = np.arange(nblocks, dtype=np.uint64)
    c1 = np.full(nblocks, c1, dtype=np.uint64)
    c2 = np.full(nblocks, c2, dtype=np.uint64)
    c3 = np.full(nblocks, c3, dtype=np.uint64)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = ((p1 >> s32) ^ c1 ^ np.uint64(k0)) & m
        n2 = ((p0 >> s32) ^ c3 ^ np.uint64(k1)) & m
        c0, c1, c2, c3 = n0, p1 & m, n2, p0 & m
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return np.stack([c0, c1, c2, c3], axis=1)


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & MASK, seed >> 32


def u53(w0, w1):
    """53-bit uniform in [0,1) from two words (same construction as numpy's random_double)."""
    return ((int(w0) >> 5) * 67108864 + (int(w1) >> 6)) * TWO_M53


class Stream:
    """Per-(seed, chain, iteration) view of the counter space with running call numbers."""

    def __init__(self, seed, chain, iteration):
        self.k0, self.k1 = split_seed(seed)
        self.chain = int(chain) & MASK
        self.iteration = int(iteration) & MASK
        self.calls = [0] * 8

    def _next(self, stream):
        n = self.calls[stream]
        self.calls[stream] = n + 1
        return ((n << 3) | stream) & MASK

    def words(self, stream, nwords):
        c1 = self._next(stream)
        nb = (nwords + 3) // 4
        if nb == 1:
            w = np.array(philox4x32(0, c1, self.iteration, self.chain, self.k0, self.k1), dtype=np.uint64)
        else:
            w = philox4x32_blocks(nb, c1, self.iteration, self.chain, self.k0, self.k1).reshape(-1)
        return w[:nwords]

    def uniform53(self, stream):
        w = self.words(stream, 2)
        return u53(w[0], w[1])

    def uniform32_vec(self, stream, n):
        return self.words(stream, n).astype(np.float64) * TWO_M32

    def normal_vec(self, n):
        """Box-Muller on 32-bit words: block b=(w0,w1,w2,w3) -> elements 4b..4b+3 =
        r(w0)cos(t(w1)), r(w0)sin(t(w1)), r(w2)cos(t(w3)), r(w2)sin(t(w3)),
        r(w)=sqrt(-2 ln((w+1)2^-32)), t(w)=2 pi w 2^-32."""
        nb = (n + 3) // 4
        w = self.words(ST_NORMAL, 4 * nb).astype(np.float64).reshape(nb, 2, 2)
        r = np.sqrt(-2.0 * np.log((w[:, :, 0] + 1.0) * TWO_M32))
        t = (2.0 * np.pi) * (w[:, :, 1] * TWO_M32)
        out = np.stack([r * np.cos(t), r * np.sin(t)], axis=2).reshape(-1)
        return out[:n]

    def sample(self, M, n):
        """n distinct integers from range(M): r_j = (w_j*(M-j))>>32, then shifted past the
        already chosen values taken in ascending order."""
        w = self.words(ST_SAMPLE, n)
        chosen = []
        for j in range(n):
            r = (int(w[j]) * (M - j)) >> 32
            for prev in sorted(chosen):
                if r >= prev:
                    r += 1
            chosen.append(r)
        return chosen

    def multinomial_index(self, p):
        u = self.uniform53(ST_MULTINOMIAL)
        acc = 0.0
        for j, pj in enumerate(p):
            acc = acc + float(pj)
            if u < acc:
                return j
        return len(p) - 1

    def randint(self, n):
        w = self.words(ST_RANDINT, 1)
        return (int(w[0]) * n) >> 32
