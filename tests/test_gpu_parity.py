"""Parity of the CUDA path (through the C ABI, via pydream_b200.engine) against
(a) the golden vectors written by the unmodified reference and (b) the C oracle on larger seeded
inputs.  Integer decisions must match bit for bit; log-posteriors within 1e-12 * max(1, |logp|)
(north_star tolerance, relative because ulp(1e4) alone is 1.8e-12)."""
import numpy as np
import pytest

from golden_util import golden_cases, load_case, make_target, prior_arrays, sampler_kwargs, decode_decisions, logp_tol

pytestmark = pytest.mark.gpu


def _engine(meta, z, **over):
    from pydream_b200.engine import DreamEngine
    d = meta['target']['d']
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    kw = sampler_kwargs(meta)
    kw.update(over)
    return DreamEngine(d, meta['N'], z['history'], z['starts'], tgt, pk, pa, pb, seed=meta['seed'], **kw)


def _run(eng, T):
    trace, logp, dec = eng.run(T)
    d = eng.d
    return (trace[:, :, :d].permute(1, 0, 2).contiguous().cpu().numpy(), logp.t().contiguous().cpu().numpy(),
            dec.t().contiguous().cpu().numpy().astype(np.uint32))


@pytest.mark.parametrize('name', golden_cases())
def test_cuda_matches_reference_golden(name):
    meta, z = load_case(name)
    eng = _engine(meta, z)
    states, logp, dec = _run(eng, meta['T'])
    dd = decode_decisions(dec)
    np.testing.assert_array_equal(dd['changed'], z['accept'])
    mn = z['multinomial']
    kw = sampler_kwargs(meta)
    col = 0
    if kw['snooker'] != 0:
        np.testing.assert_array_equal(dd['snooker'], (mn[:, :, 0] == 0).astype(int))
        col = 1
    np.testing.assert_array_equal(dd['cr'], mn[:, :, col])
    np.testing.assert_array_equal(dd['lvl'], mn[:, :, col + 1])
    k = kw['multitry']
    if k > 1:
        total = (mn >= 0).sum(axis=2)
        sel_col = np.where(dd['snooker'].astype(bool), total - 2, total - k)
        np.testing.assert_array_equal(dd['sel'], np.take_along_axis(mn, sel_col[:, :, None], axis=2)[:, :, 0])
    ref_logp = z['log_like'] + z['log_prior']
    err = np.abs(logp - ref_logp)
    assert np.all(err <= logp_tol(ref_logp)), err.max()
    np.testing.assert_allclose(states, z['states'], rtol=1e-10, atol=1e-11)
    hf = eng.history_flat()
    assert hf.shape == z['history_final'].shape
    np.testing.assert_allclose(hf, z['history_final'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.cr_probs.cpu().numpy(), z['cr_probs'][-1], rtol=1e-10)
    np.testing.assert_allclose(eng.gamma_probs.cpu().numpy(), z['gamma_probs'][-1], rtol=1e-10)
    np.testing.assert_array_equal(eng.ncr_updates.cpu().numpy(), z['ncr_updates'])


CASES = [
    # name, d, N, T, target, kwargs
    ('c2_gauss100', 100, 256, 60, 'gaussian', dict(snooker=.1, history_thin=10)),
    ('c2_gauss100_de', 100, 200, 45, 'gaussian', dict(snooker=0., history_thin=10)),
    ('c3_mix10_mt5', 10, 512, 40, 'mixture', dict(multitry=5, snooker=.1, history_thin=10)),
    ('c4_banana200', 200, 96, 30, 'banana', dict(snooker=.1, history_thin=5)),
    ('c5_gauss50_adapt', 50, 384, 64, 'gaussian', dict(snooker=.1, history_thin=10, adapt_crossover=True, crossover_burnin=40)),
    ('gauss7_odd', 7, 33, 50, 'gaussian', dict(snooker=.2, history_thin=3, DEpairs=2, multitry=3)),
    ('gauss300', 300, 40, 12, 'gaussian', dict(snooker=.1, history_thin=4)),
    ('banana530_r8', 530, 24, 10, 'banana', dict(snooker=.3, history_thin=2)),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_c_oracle(case):
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    name, d, N, T, tkind, kw = case
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    tgt = make_target(dict(kind=tkind, d=d))
    nseed = 2 * N * kw.get('DEpairs', 1) + 17
    hist = rng.uniform(-5, 15, size=(nseed, d)) if tkind != 'mixture' else rng.normal(size=(nseed, d))
    starts = hist[:N].copy()
    okw = dict(adapt_crossover=False, adapt_gamma=False, crossover_burnin=0)
    okw.update(kw)
    orc = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=5, nthreads=4, **okw)
    ref = orc.run(T)
    eng = DreamEngine(d, N, hist, starts, tgt, seed=5, **okw)
    states, logp, dec = _run(eng, T)
    np.testing.assert_array_equal(dec, ref['decisions'])
    err = np.abs(logp - ref['logp'])
    assert np.all(err <= logp_tol(ref['logp'])), (err / logp_tol(ref['logp'])).max()
    np.testing.assert_allclose(states, ref['states'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.history_flat(), orc.history_flat, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.cr_probs.cpu().numpy(), orc.cr_probs, rtol=1e-10)


def test_gelman_rubin_kernel():
    import torch
    from oracle import c_oracle
    from pydream_b200.convergence import Gelman_Rubin
    rng = np.random.default_rng(3)
    chains = [rng.normal(size=(501, 7)) * (1 + .1 * c) + .05 * c for c in range(5)]
    got = Gelman_Rubin(chains)
    ref = c_oracle.gelman_rubin(np.stack(chains))
    np.testing.assert_allclose(got, ref, rtol=1e-12)
    # definition check against numpy (pydream/convergence.py:3-20 restated)
    arr = np.stack(chains)
    nb = 501 // 2
    W = np.mean([np.var(c[nb:], axis=0) for c in arr], axis=0)
    B = np.var([np.mean(c[nb:], axis=0) for c in arr], axis=0)
    np.testing.assert_allclose(got, np.sqrt((W * (1 - 1. / 501) + B) / W), rtol=1e-12)


def test_fused_window_equals_single_steps():
    """Fusing up to history_thin iterations per launch must not change anything."""
    meta, z = load_case('gauss30_snooker')
    a = _run(_engine(meta, z), meta['T'])
    eng = _engine(meta, z)
    parts = [_run(eng, 1) for _ in range(meta['T'])]
    for i in range(3):
        np.testing.assert_array_equal(a[i], np.concatenate([p[i] for p in parts], axis=0))


def test_gauss_kernel_equals_generic_kernel():
    """The CTA-synchronous dense-Gaussian kernel (TMA-staged rows, shared quadratic forms) and the generic
    lane-group kernel consume the same random streams: identical decisions, logp within tolerance."""
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(77)
    d, N, T = 100, 1031, 43      # N not a multiple of the chains-per-CTA tile
    tgt = make_target(dict(kind='gaussian', d=d))
    hist = rng.uniform(-5, 15, size=(2 * N + 5, d))
    kw = dict(seed=9, snooker=.15, history_thin=7, DEpairs=2)
    a = _run(DreamEngine(d, N, hist, hist[:N], tgt, **kw), T)
    b = _run(DreamEngine(d, N, hist, hist[:N], tgt, generic_kernel=True, **kw), T)
    np.testing.assert_array_equal(a[2], b[2])
    assert np.all(np.abs(a[1] - b[1]) <= logp_tol(b[1]))
    np.testing.assert_allclose(a[0], b[0], rtol=1e-10, atol=1e-11)


def test_window_kernel_equals_generic_kernel():
    """The dense-Gaussian window kernel (incremental quadratic form, batched draws, TMA-staged rows) consumes the
    same random streams as the generic kernel: identical decisions, logp within tolerance, for any split of the
    run into launches."""
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(78)
    d, N, T = 100, 1031, 47      # N not a multiple of the chains-per-CTA tile
    tgt = make_target(dict(kind='gaussian', d=d))
    hist = rng.uniform(-5, 15, size=(2 * N + 5, d))
    kw = dict(seed=10, snooker=.15, history_thin=7)
    a = _run(DreamEngine(d, N, hist, hist[:N], tgt, **kw), T)
    b = _run(DreamEngine(d, N, hist, hist[:N], tgt, generic_kernel=True, **kw), T)
    c = _run(DreamEngine(d, N, hist, hist[:N], tgt, window_kernel=False, **kw), T)
    for other in (b, c):
        np.testing.assert_array_equal(a[2], other[2])
        assert np.all(np.abs(a[1] - other[1]) <= logp_tol(other[1])), (np.abs(a[1] - other[1]) / logp_tol(other[1])).max()
        np.testing.assert_allclose(a[0], other[0], rtol=1e-10, atol=1e-11)
    eng = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    parts = [_run(eng, n) for n in (1, 2, 3, 11, 30)]
    for i in range(3):
        np.testing.assert_array_equal(a[i], np.concatenate([p[i] for p in parts], axis=0))


@pytest.mark.parametrize('d', [68, 97, 128])
def test_window_kernel_dimensions(d):
    """Row strides that are / are not multiples of 8 doubles, odd dimensions, the largest supported row."""
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(d)
    N, T = 50, 33
    tgt = make_target(dict(kind='gaussian', d=d))
    hist = rng.uniform(-5, 15, size=(2 * N + 3, d))
    kw = dict(snooker=.2, history_thin=4)
    ref = c_oracle.OracleSampler(d, N, hist, hist[:N], tgt.kind, tgt.table(), seed=3, nthreads=4, **kw).run(T)
    states, logp, dec = _run(DreamEngine(d, N, hist, hist[:N], tgt, seed=3, **kw), T)
    np.testing.assert_array_equal(dec, ref['decisions'])
    assert np.all(np.abs(logp - ref['logp']) <= logp_tol(ref['logp']))
    np.testing.assert_allclose(states, ref['states'], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize('kw', [
    dict(nCR=5, gamma_levels=3, adapt_gamma=True, adapt_crossover=True, crossover_burnin=30, snooker=.15, history_thin=6),
    dict(nCR=16, gamma_levels=8, snooker=.1, history_thin=1, p_gamma_unity=.5),
    dict(nCR=1, snooker=0., history_thin=25, p_gamma_unity=0., lamb=.2, zeta=1e-6),
], ids=['adapt_cr_and_gamma', 'max_tables_thin1', 'no_snooker_long_window'])
def test_window_kernel_option_space(kw):
    """Corners of the option space on the window kernel against the generic kernel (same draws): many CR values and
    gamma levels with both adaptations during burn-in, a window of one iteration, a window longer than a batch."""
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(91)
    d, N, T = 80, 77, 52
    tgt = make_target(dict(kind='gaussian', d=d))
    hist = rng.uniform(-5, 15, size=(2 * N + 5, d))
    a = _run(DreamEngine(d, N, hist, hist[:N], tgt, seed=12, **kw), T)
    b = _run(DreamEngine(d, N, hist, hist[:N], tgt, seed=12, generic_kernel=True, **kw), T)
    np.testing.assert_array_equal(a[2], b[2])
    assert np.all(np.abs(a[1] - b[1]) <= logp_tol(b[1])), (np.abs(a[1] - b[1]) / logp_tol(b[1])).max()
    np.testing.assert_allclose(a[0], b[0], rtol=1e-10, atol=1e-11)


def test_rng_normals_match_contract_bit_for_bit():
    """The kernels' float32 Box-Muller (csrc/dreamzs_common.cuh normal_pair32) against the numpy statement of the RNG
    contract (oracle/philox.py): identical bits for 16384 blocks of several (chain, iteration, call) streams."""
    import ctypes as C
    import torch
    from oracle import philox as px
    from pydream_b200 import _cabi
    lib = _cabi.load()
    nb = 16384
    for seed, chain, it, call in ((0, 0, 0, 0), (0x123456789ABCDEF, 4095, 77, 3), (2 ** 64 - 1, 2 ** 32 - 2, 10 ** 6, 1)):
        out = torch.empty(4 * nb, dtype=torch.float32, device='cuda')
        _cabi.check(lib.dreamzs_rng_normals(C.c_uint64(seed), chain, it, call, nb, C.c_void_p(out.data_ptr()), None), 'dreamzs_rng_normals')
        k0, k1 = px.split_seed(seed)
        w = px.philox4x32_blocks(nb, (call << 3) | px.ST_NORMAL, it, chain, k0, k1).astype(np.uint32).reshape(nb, 2, 2)
        n0, n1 = px.normal_pairs32(w[:, :, 0], w[:, :, 1])
        ref = np.stack([n0, n1], axis=2).reshape(-1)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize('shape', [(50, 333, 41, 10), (100, 150, 57, 25), (24, 700, 33, 10), (128, 40, 26, 4), (9, 900, 21, 1)],
                         ids=['d50', 'd100_thin25', 'd24', 'd128', 'd9_thin1'])
def test_whitened_window_kernel_shapes(shape):
    """The whitened window kernel (DMMA products, flattened draws) over its shape space -- 8 / 16 / 32 lanes per chain,
    windows longer than a batch, one-iteration windows, chain counts that do not fill the last CTA -- against the C
    oracle and against the same engine with the kernel switched off (generic path, same draws)."""
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    d, N, T, thin = shape
    rng = np.random.default_rng(d * 1000 + N)
    tgt = make_target(dict(kind='gaussian', d=d))
    hist = rng.uniform(-5, 15, size=(2 * N + 3, d))
    kw = dict(snooker=.15, history_thin=thin)
    ref = c_oracle.OracleSampler(d, N, hist, hist[:N], tgt.kind, tgt.table(), seed=8, nthreads=8, **kw).run(T)
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=8, **kw)
    assert eng.gauss_L is not None
    states, logp, dec = _run(eng, T)
    np.testing.assert_array_equal(dec, ref['decisions'])
    assert np.all(np.abs(logp - ref['logp']) <= logp_tol(ref['logp'])), (np.abs(logp - ref['logp']) / logp_tol(ref['logp'])).max()
    np.testing.assert_allclose(states, ref['states'], rtol=1e-10, atol=1e-11)
    gen = _run(DreamEngine(d, N, hist, hist[:N], tgt, seed=8, generic_kernel=True, **kw), T)
    np.testing.assert_array_equal(dec, gen[2])
    # any split of the run into launches gives the same trajectory
    eng2 = DreamEngine(d, N, hist, hist[:N], tgt, seed=8, **kw)
    parts = [_run(eng2, n) for n in (1, 2, 7, T - 10)]
    np.testing.assert_array_equal(dec, np.concatenate([p[2] for p in parts], axis=0))
    np.testing.assert_array_equal(logp, np.concatenate([p[1] for p in parts], axis=0))
