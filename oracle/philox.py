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


# ---------------------------------------------------------------------------------------------------------------------
# Normal variates of the contract (stream 2).  zeta of Dream.py:694 is N(0, 1e-12): a jitter ~12 orders of magnitude
# below the states it is added to, so float32 resolution is ample; what the contract needs is that every
# implementation (this file, oracle/dreamzs_oracle.c, the CUDA kernels) produces THE SAME BITS.  The recipe below
# therefore uses only operations that IEEE-754 rounds identically everywhere -- float32 add, multiply, fused
# multiply-add and square root, round to nearest even -- on fixed polynomial coefficients (Cephes logf / sinf / cosf):
#
#   pair(w0, w1):  a = (w0 >> 8) + 1                      in [1, 2^24]        (u = a 2^-24 in (0, 1])
#                  a = m 2^E, m in [1, 2);  m > 1.41421356f: m <- m/2, E <- E+1;  t = m - 1
#                  log m = t + (t (t^2 P(t)) - 0.5 t^2),  P = Horner of LOG_P with fma
#                  L = fma(E - 24, ln 2, log m);  r = sqrt(max(-2 L, 0))
#                  v = (w1 >> 8) 2^-24 in [0, 1):  k = round(4 v) (0..4), g = v - k/4 in [-1/8, 1/8), phi = g * 2 pi
#                  z = phi^2;  s = fma(phi z, SIN_P(z), phi);  c = fma(z z, COS_P(z), fma(-0.5, z, 1))
#                  (cos, sin)(2 pi v) = quadrant rotation k & 3 of (c, s):  (c, s), (-s, c), (-c, -s), (s, -c)
#                  pair = (r * cos, r * sin)
# numpy has no fused multiply-add: `fma32` evaluates a*b exactly in float64, adds c with round-to-odd (TwoSum error
# term) and rounds once to float32, which equals the IEEE float32 fma for all inputs used here.
F32 = np.float32
LOG_P = [F32(x) for x in (7.0376836292E-2, -1.1514610310E-1, 1.1676998740E-1, -1.2420140846E-1, 1.4249322787E-1,
                          -1.6668057665E-1, 2.0000714765E-1, -2.4999993993E-1, 3.3333331174E-1)]
SIN_P = [F32(x) for x in (-1.9515295891E-4, 8.3321608736E-3, -1.6666654611E-1)]
COS_P = [F32(x) for x in (2.443315711809948E-5, -1.388731625493765E-3, 4.166664568298827E-2)]
LN2_F = F32(0.6931471805599453)
SQRT2_F = F32(1.41421356)
TWO_PI_F = F32(6.283185307179586)


def fma32(a, b, c):
    """IEEE float32 fused multiply-add on numpy arrays (exact product in float64, round-to-odd sum, one rounding)."""
    a, b, c = (np.asarray(x, dtype=np.float32).astype(np.float64) for x in (a, b, c))
    p = a * b                                   # exact: 24 + 24 significant bits
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)             # TwoSum: p + c == s + err exactly
    si = np.atleast_1d(s).view(np.int64).copy()
    up = (np.atleast_1d(err) > 0) == (np.atleast_1d(s) > 0)          # the exact sum is larger in magnitude than s
    fix = (np.atleast_1d(err) != 0) & ((si & 1) == 0)
    si = np.where(fix, si + np.where(up, 1, -1), si)
    return si.view(np.float64).reshape(np.shape(s)).astype(np.float32)


def _horner32(coefs, x):
    acc = np.full(np.shape(x), coefs[0], dtype=np.float32)
    for c in coefs[1:]:
        acc = fma32(acc, x, c)
    return acc


def normal_pairs32(w0, w1):
    """Two standard normals (float32) per pair of 32-bit words; see the recipe above."""
    w0 = np.asarray(w0, dtype=np.uint32)
    w1 = np.asarray(w1, dtype=np.uint32)
    a = ((w0 >> np.uint32(8)) + np.uint32(1)).astype(np.float32)            # exact
    bits = a.view(np.uint32)
    E = (bits >> np.uint32(23)).astype(np.int32) - 127
    m = ((bits & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)).view(np.float32)
    big = m > SQRT2_F
    m = np.where(big, m * F32(0.5), m).astype(np.float32)
    E = E + big.astype(np.int32)
    t = (m - F32(1.0)).astype(np.float32)
    z = (t * t).astype(np.float32)
    y = (t * (z * _horner32(LOG_P, t)).astype(np.float32)).astype(np.float32)
    y = fma32(F32(-0.5), z, y)
    logm = (t + y).astype(np.float32)
    L = fma32((E - 24).astype(np.float32), LN2_F, logm)
    r = np.sqrt(np.maximum((F32(-2.0) * L).astype(np.float32), F32(0.0))).astype(np.float32)
    w24 = (w1 >> np.uint32(8)).astype(np.int64)
    k = (w24 + (1 << 21)) >> 22
    g = ((w24 - (k << 22)).astype(np.float32) * F32(2.0 ** -24)).astype(np.float32)   # exact
    phi = (g * TWO_PI_F).astype(np.float32)
    zz = (phi * phi).astype(np.float32)
    s = fma32((phi * zz).astype(np.float32), _horner32(SIN_P, zz), phi)
    c = fma32((zz * zz).astype(np.float32), _horner32(COS_P, zz), fma32(F32(-0.5), zz, F32(1.0)))
    q = k & 3
    cs = np.where(q == 0, c, np.where(q == 1, -s, np.where(q == 2, -c, s))).astype(np.float32)
    sn = np.where(q == 0, s, np.where(q == 1, c, np.where(q == 2, -s, -c))).astype(np.float32)
    return (r * cs).astype(np.float32), (r * sn).astype(np.float32)


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
        """Box-Muller in float32 on 24-bit uniforms (`normal_pairs32`): block b=(w0,w1,w2,w3) -> elements 4b..4b+3 =
        pair(w0,w1), pair(w2,w3); returned as float64 (exact widening)."""
        nb = (n + 3) // 4
        w = self.words(ST_NORMAL, 4 * nb).astype(np.uint32).reshape(nb, 2, 2)
        n0, n1 = normal_pairs32(w[:, :, 0], w[:, :, 1])
        out = np.stack([n0, n1], axis=2).reshape(-1).astype(np.float64)
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
