"""The reference's frequency tests (pydream/tests/test_dream.py:54-149, 202-305), restated on the decision words and
traces of the C oracle: the draws served from the Philox contract must have the distributions the reference's tests
demand of np.random.  (The CUDA path takes bit-identical decisions: tests/test_gpu_parity.py.)"""
import numpy as np
import pytest

from oracle import c_oracle
from pydream_b200 import targets
from golden_util import decode_decisions


def _run(d=4, N=200, T=60, target=None, **kw):
    rng = np.random.default_rng(11)
    tgt = target or targets.Constant(d, 0.0)           # flat prior + constant likelihood: every proposal is accepted
    hist = rng.normal(size=(2 * N * kw.get('DEpairs', 1) + 9, d))
    s = c_oracle.OracleSampler(d, N, hist, hist[:N], tgt.kind, tgt.table(), seed=77, nthreads=4, **kw)
    out = s.run(T)
    return s, out, decode_decisions(out['decisions'])


def test_snooker_fraction():
    """test_snooker_fraction (:86-98): P(snooker) = .1 within .05."""
    _, _, dd = _run(snooker=.1)
    assert abs(dd['snooker'].mean() - .1) < .02
    _, _, dd = _run(snooker=0.)
    assert dd['snooker'].sum() == 0


def test_CR_fraction():
    """test_CR_fraction (:100-123): crossover values are drawn with the given probabilities."""
    probs = [.1, .6, .3]
    _, _, dd = _run(nCR=3, cr_probs=probs, snooker=0.)
    freq = np.bincount(dd['cr'].reshape(-1), minlength=3) / dd['cr'].size
    np.testing.assert_allclose(freq, probs, atol=.02)


def test_DEpair_selec():
    """test_DEpair_selec (:125-149): the number of DE pairs is uniform over 1..DEpairs."""
    _, _, dd = _run(d=6, DEpairs=3, snooker=0.)
    freq = np.bincount(dd['delta'].reshape(-1), minlength=4)[1:] / dd['delta'].size
    np.testing.assert_allclose(freq, [1 / 3.] * 3, atol=.02)


def test_gamma_unityfraction():
    """test_gamma_unityfraction (:54-66): gamma = 1 with probability p_gamma_unity for DE moves."""
    _, _, dd = _run(snooker=0., p_gamma_unity=.2)
    assert abs(dd['gamma_one'].mean() - .2) < .02


def test_gamma_level_fraction():
    probs = [.5, .25, .125, .125]
    _, _, dd = _run(snooker=0., gamma_levels=4, gamma_probs=probs)
    freq = np.bincount(dd['lvl'].reshape(-1), minlength=4) / dd['lvl'].size
    np.testing.assert_allclose(freq, probs, atol=.02)


@pytest.mark.parametrize('m', [1, 2, 3])
def test_proposal_generation_nosnooker_CR(m):
    """test_proposal_generation_nosnooker_CR1 / CR33 / CR66 (:202-305): with crossover value CR a dimension keeps the
    current value with probability 1 - CR.  Every proposal is accepted here (constant likelihood, flat prior), so the
    trace shows the proposals themselves."""
    d, N, T = 4, 200, 40
    probs = [0., 0., 0.]
    probs[m - 1] = 1.
    s, out, dd = _run(d=d, N=N, T=T, nCR=3, cr_probs=probs, snooker=0.)
    assert np.all(dd['cr'] == m - 1)
    states = out['states']                                   # (T, N, d)
    changed = states[1:] != states[:-1]
    CR = m / 3.
    assert abs(changed.mean() - CR) < .03
    # a proposal with no dimension selected equals the current point and counts as a rejection (Dream.py:336)
    none_sel = ~changed.any(axis=2)
    np.testing.assert_array_equal(dd['changed'][1:][none_sel], 0)
    assert abs(none_sel.mean() - (1 - CR) ** d) < .02


def test_chain_sampling_rows_are_archive_rows():
    """test_chain_sampling_* (:160-200): sampled rows are rows of the archive, distinct within a call."""
    rng = np.random.default_rng(5)
    d, N = 3, 40
    hist = rng.normal(size=(120, d))
    tgt = targets.Constant(d, 0.0)
    s = c_oracle.OracleSampler(d, N, hist, hist[:N], tgt.kind, tgt.table(), seed=3, snooker=0., DEpairs=2, history_thin=1000)
    out = s.run(30, rows_dbg_n=8)
    rows = out['rows']                                       # (T, N, 8), -1 padded
    used = rows[rows >= 0]
    assert used.min() >= 0 and used.max() < 120 + N         # the seed rows + the one append of iteration 0
    for t in range(rows.shape[0]):
        for c in range(N):
            r = rows[t, c][rows[t, c] >= 0]
            assert len(r) in (2, 4) and len(set(r.tolist())) == len(r)
    # every archive row is reachable, roughly uniformly
    cnt = np.bincount(used, minlength=160)[:120]
    assert cnt.min() > 0
