"""Pins the C restatement (oracle/dreamzs_oracle.c) against the golden vectors written by the
unmodified reference under lock-step execution (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import c_oracle
from golden_util import golden_cases, load_case, make_target, prior_arrays, sampler_kwargs, decode_decisions, logp_tol


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
