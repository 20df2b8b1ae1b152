"""Pins the C restatement (oracle/dreamzs_oracle.c) against the golden vectors written by the
unmodified reference under lock-step execution (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import c_oracle
from golden_util import golden_cases, golden_pt_cases, load_case, make_target, prior_arrays, sampler_kwargs, decode_decisions, logp_tol


def run_oracle(meta, z, nthreads=1):
    d = meta['target']['d']
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    s = c_oracle.OracleSampler(d, meta['N'], z['history'], z['starts'], tgt.kind, tgt.table(), seed=meta['seed'],
                               prior_kind=pk, prior_a=pa, prior_b=pb, nthreads=nthreads, **sampler_kwargs(meta))
    out = s.run(meta['T'], rows_dbg_n=32)
    return s, out


@pytest.mark.parametrize('name', golden_cases())
def test_oracle_matches_reference(name):
    meta, z = load_case(name)
    _check_against_reference(meta, z)


def _check_against_reference(meta, z):
    s, out = run_oracle(meta, z)
    dec = decode_decisions(out['decisions'])
    # integer decisions: bit-exact
    np.testing.assert_array_equal(dec['changed'], z['accept'])
    k = sampler_kwargs(meta)['multitry']
    nrow = min(32, z['rows'].shape[2])
    np.testing.assert_array_equal(out['rows'][:, :, :nrow], z['rows'][:, :, :nrow])
    mn = z['multinomial']
    has_snk = sampler_kwargs(meta)['snooker'] != 0
    col = 0
    if has_snk:
        np.testing.assert_array_equal(dec['snooker'], (mn[:, :, 0] == 0).astype(int))
        col = 1
    np.testing.assert_array_equal(dec['cr'], mn[:, :, col])
    np.testing.assert_array_equal(dec['lvl'], mn[:, :, col + 1])
    if k > 1:
        # multinomial order: [snooker] CR level, then k (DE) or 1 (snooker) gamma-unity draws per generated
        # proposal batch (repeated when a batch is regenerated, Dream.py:282-289), the selection, then k-1 (DE) or
        # 1 (snooker) gamma-unity draws of the reference batch
        snk = dec['snooker'].astype(bool)
        total = (mn >= 0).sum(axis=2)
        sel_col = np.where(snk, total - 2, total - k)
        sel = np.take_along_axis(mn, sel_col[:, :, None], axis=2)[:, :, 0]
        np.testing.assert_array_equal(dec['sel'], sel)
    # floating point: states and log-posteriors
    ref_logp = z['log_like'] + z['log_prior']
    assert np.all(np.abs(out['logp'] - ref_logp) <= logp_tol(ref_logp)), np.abs(out['logp'] - ref_logp).max()
    # snooker projections (sum of cancelling terms divided by a BLAS dot) amplify 1-ulp differences
    np.testing.assert_allclose(out['states'], z['states'], rtol=1e-10, atol=1e-11)
    # archive layout and content (record_history, row-major flat float64)
    hf = s.history_flat
    assert hf.shape == z['history_final'].shape
    np.testing.assert_allclose(hf, z['history_final'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(s.cr_probs, z['cr_probs'][-1], rtol=1e-12)
    np.testing.assert_allclose(s.gamma_probs, z['gamma_probs'][-1], rtol=1e-12)
    np.testing.assert_allclose(s.delta_m, z['delta_m'], rtol=1e-12)
    np.testing.assert_array_equal(s.ncr_updates, z['ncr_updates'])


@pytest.mark.parametrize('name', golden_pt_cases())
def test_oracle_matches_reference_tempering(name):
    """Parallel tempering: the C restatement against the reference's own _sample_dream_pt driver
    (pydream/core.py:131-236) run on the lock-step pool."""
    meta, z = load_case(name)
    d = meta['target']['d']
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    s = c_oracle.OracleSampler(d, meta['N'], z['history'], z['starts'], tgt.kind, tgt.table(), seed=meta['seed'],
                               prior_kind=pk, prior_a=pa, prior_b=pb, **sampler_kwargs(meta))
    np.testing.assert_array_equal(c_oracle.temperature_ladder(meta['N']), z['T'])
    out = s.run_pt(meta['T'])
    assert out['sampled_params'].shape == z['sampled_params'].shape             # (N, 2 niter, d)
    np.testing.assert_array_equal(out['swaps'][:, :2], z['pairs'])                # the chains proposed for a swap
    ref_swapped = np.any(z['sampled_params'][:, 0::2] != z['sampled_params'][:, 1::2], axis=(0, 2))
    got_swapped = np.any(out['sampled_params'][:, 0::2] != out['sampled_params'][:, 1::2], axis=(0, 2))
    np.testing.assert_array_equal(got_swapped, ref_swapped)
    assert ref_swapped.sum() > 3 and (~ref_swapped).sum() > 3
    # accept / reject sequence of the steps: state changed between the post-swap row of t-1 and the step row of t
    ref_sp, got_sp = z['sampled_params'], out['sampled_params']
    ref_acc = np.any(ref_sp[:, 2::2] != ref_sp[:, 1:-1:2], axis=2)
    np.testing.assert_array_equal(np.any(got_sp[:, 2::2] != got_sp[:, 1:-1:2], axis=2), ref_acc)
    np.testing.assert_array_equal((out['decisions'][:, 2::2] & 1).astype(bool), ref_acc)
    np.testing.assert_allclose(got_sp, ref_sp, rtol=1e-10, atol=1e-11)
    ref_lp = z['log_ps'][:, :, 0]
    assert np.all(np.abs(out['log_ps'] - ref_lp) <= logp_tol(ref_lp)), np.abs(out['log_ps'] - ref_lp).max()
    hf = s.history_flat
    assert hf.shape == z['history_final'].shape
    np.testing.assert_allclose(hf, z['history_final'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(s.cr_probs, z['cr_probs'][-1], rtol=1e-12)


def _random_case(i):
    """A small random configuration inside what the reference itself supports (niterations a multiple of
    history_thin: its archive overflows otherwise, Dream.py:936; no multi-try in one dimension)."""
    rng = np.random.default_rng(1000 + i)
    kind = ['gaussian', 'banana', 'mixture', 'sumshift'][i % 4]
    d = int(rng.integers(2, 13))
    DEpairs = int(rng.integers(1, 3))
    N = int(rng.integers(2 * DEpairs + 1, 2 * DEpairs + 5))
    thin = int(rng.integers(1, 5))
    T = thin * int(rng.integers(4, 9))
    kw = dict(snooker=float(rng.choice([0., .1, .5])), history_thin=thin, DEpairs=DEpairs,
              multitry=[False, False, 3, 5][int(rng.integers(0, 4))],   # (1 == True means 5 in the reference)
               nCR=int(rng.integers(1, min(d, 4) + 1)), adapt_crossover=False,
              crossover_burnin=3, p_gamma_unity=float(rng.choice([.2, .5])), gamma_levels=int(rng.integers(1, 4)))
    if kind == 'sumshift':
        prior = dict(kind='uniform', loc=[-4.] * d, scale=[9.] * d)
        hist = -4. + 9. * rng.uniform(size=(2 * DEpairs * N + 6, d))
        kw['lamb'] = .4
    else:
        prior = dict(kind='flat', d=d)
        hist = rng.normal(size=(2 * DEpairs * N + 6, d)) * (3. if kind != 'mixture' else 1.)
    return dict(target=dict(kind=kind, d=d), prior=prior, N=N, T=T, seed=500 + i, kw=kw), hist


@pytest.mark.parametrize('i', range(16))
def test_oracle_matches_live_reference(i):
    """Fresh random cases through the UNMODIFIED reference (lock-step harness) and the C restatement, side by side.
    Runs where /root/reference exists (the build container); the committed golden vectors cover the GPU box."""
    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip('reference checkout not present')
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden as G
    meta, hist = _random_case(i)
    starts = hist[:meta['N']].copy()
    z = H.run_lockstep(G.make_params(meta['prior']), make_target(meta['target']), meta['N'], meta['T'], starts, hist,
                       seed=meta['seed'], **meta['kw'])
    z = dict(z, history=hist, starts=starts)
    _check_against_reference(meta, z)


@pytest.mark.parametrize('i', range(6))
def test_oracle_matches_live_reference_tempering(i):
    """The same for parallel tempering: the reference's _sample_dream_pt on the lock-step pool vs dreamzs_oracle_run_pt."""
    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip('reference checkout not present')
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    import make_golden as G
    meta, hist = _random_case(100 + i)
    d, N = meta['target']['d'], meta['N']
    starts = hist[:N].copy()
    z = H.run_lockstep_pt(G.make_params(meta['prior']), make_target(meta['target']), N, meta['T'], starts, hist,
                          seed=meta['seed'], **meta['kw'])
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    s = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=meta['seed'], prior_kind=pk, prior_a=pa,
                               prior_b=pb, **sampler_kwargs(meta))
    out = s.run_pt(meta['T'])
    np.testing.assert_array_equal(out['swaps'][:, :2], z['pairs'])
    ref_sp, got_sp = z['sampled_params'], out['sampled_params']
    np.testing.assert_array_equal(np.any(got_sp[:, 0::2] != got_sp[:, 1::2], axis=(0, 2)),
                                  np.any(ref_sp[:, 0::2] != ref_sp[:, 1::2], axis=(0, 2)))
    np.testing.assert_array_equal(np.any(got_sp[:, 2::2] != got_sp[:, 1:-1:2], axis=2),
                                  np.any(ref_sp[:, 2::2] != ref_sp[:, 1:-1:2], axis=2))
    np.testing.assert_allclose(got_sp, ref_sp, rtol=1e-10, atol=1e-11)
    ref_lp = z['log_ps'][:, :, 0]
    assert np.all(np.abs(out['log_ps'] - ref_lp) <= logp_tol(ref_lp)), np.abs(out['log_ps'] - ref_lp).max()
    np.testing.assert_allclose(s.history_flat, z['history_final'], rtol=1e-10, atol=1e-11)


def test_oracle_threads_identical():
    meta, z = load_case('mix10_mt5')
    _, a = run_oracle(meta, z, nthreads=1)
    _, b = run_oracle(meta, z, nthreads=3)
    for k in ('states', 'logp', 'decisions'):
        np.testing.assert_array_equal(a[k], b[k])


def test_gamma_table_known_answers():
    """pydream/tests/test_dream.py:68-76: gamma = 2.38/sqrt(2*delta*d') for d'=1."""
    tab = c_oracle.gamma_table(1, 5, 4)
    for delta, val in zip(range(1, 6), [1.683, 1.19, 0.972, 0.841, 0.753]):
        assert round(abs(tab[0, delta - 1, 0] - val), 3) == 0
    tab = c_oracle.gamma_table(3, 2, 7)
    ref = np.zeros((3, 2, 7))
    dec = 1
    for lvl in range(1, 4):
        for delta in range(1, 3):
            ref[lvl - 1, delta - 1, :] = (2.38 / np.sqrt(2 * delta * np.linspace(1, 7, num=7))) / dec
        dec *= 2
    np.testing.assert_array_equal(tab, ref)


def test_gelman_rubin_matches_reference_function():
    """`Gelman_Rubin` has no test in the reference: the C restatement against the imported
    pydream.convergence.Gelman_Rubin (pydream/convergence.py:3-20) on random chains, odd and even lengths."""
    from oracle import ref_harness as H
    if not H.reference_available():
        pytest.skip('reference not present (neither /root/reference nor baseline/_ref)')
    import sys
    if H.REF_ROOT not in sys.path:
        sys.path.insert(0, H.REF_ROOT)
    from pydream.convergence import Gelman_Rubin
    rng = np.random.default_rng(11)
    for nchains, nsamples, d in ((3, 40, 1), (5, 501, 7), (8, 1000, 10), (16, 63, 4)):
        chains = [rng.normal(size=(nsamples, d)) * (1 + .2 * c) + .1 * c for c in range(nchains)]
        ref = Gelman_Rubin(chains)
        got = c_oracle.gelman_rubin(np.stack(chains))
        np.testing.assert_allclose(got, ref, rtol=1e-12)
