"""The multi-try kernels (pydream_b200/csrc/dreamzs_mtp_kernel.cuh) against the C oracle: the two-stage form (draw kernel
+ chain kernel, the default), the fused point-parallel form (two_stage=False) and the generic lane-group kernel
(generic_kernel=True) must all reproduce the oracle's decisions bit for bit -- including boundary redraws (whose
rand() call numbers depend on the points before them, Dream.py:734-791) and regenerated proposal batches
(Dream.py:278-289), which send an iteration of the two-stage form through its out-of-line path."""
import zlib

import numpy as np
import pytest

from golden_util import make_target, prior_arrays, logp_tol

pytestmark = pytest.mark.gpu

# name, d, N, T, target, prior, kwargs
CASES = [
    ('mix10_mt5', 10, 300, 40, 'mixture', dict(kind='flat'), dict(multitry=5, snooker=.1, history_thin=10)),
    ('mix10_mt8_snk', 10, 64, 30, 'mixture', dict(kind='flat'), dict(multitry=8, snooker=.5, history_thin=4)),
    ('mix3_mt3_pairs3', 3, 40, 36, 'mixture', dict(kind='flat'), dict(multitry=3, snooker=.2, history_thin=3, DEpairs=3)),
    ('gauss20_mt3', 20, 50, 30, 'gaussian', dict(kind='flat'), dict(multitry=3, snooker=.3, history_thin=5, gamma_levels=3)),
    ('gauss30_mt4_pairs2', 30, 33, 24, 'gaussian', dict(kind='flat'), dict(multitry=4, snooker=.1, history_thin=6, DEpairs=2)),
    ('banana16_mt4', 16, 48, 30, 'banana', dict(kind='flat'), dict(multitry=4, snooker=.1, history_thin=5, p_gamma_unity=.5)),
    # uniform priors + hard boundaries: reflections and redraws in nearly every batch
    ('sum6_mt5_bounds', 6, 80, 40, 'sumshift', dict(kind='uniform', loc=[-.5] * 6, scale=[1.] * 6),
     dict(multitry=5, snooker=.2, history_thin=5, zeta=1e-3, lamb=.4)),
    ('const12_mt3_bounds', 12, 60, 30, 'constant', dict(kind='uniform', loc=[0.] * 12, scale=[.7] * 12),
     dict(multitry=3, snooker=.3, history_thin=3, DEpairs=2)),
    # uniform priors WITHOUT hard boundaries: proposals leave the support, whole batches are -inf and get regenerated
    ('const4_mt3_regen', 4, 60, 40, 'constant', dict(kind='uniform', loc=[0.] * 4, scale=[.25] * 4),
     dict(multitry=3, snooker=.2, history_thin=4, hardboundaries=False, zeta=1e-3)),
    ('sum10_mt5_regen', 10, 40, 30, 'sumshift', dict(kind='uniform', loc=[-1.] * 10, scale=[.4] * 10),
     dict(multitry=5, snooker=0., history_thin=5, hardboundaries=False)),
    ('norm5_mt3', 5, 40, 30, 'sumshift', dict(kind='norm', loc=[.3] * 5, scale=[2.] * 5), dict(multitry=3, snooker=.1, history_thin=2)),
]
FORMS = [('two_stage', dict()), ('fused', dict(two_stage=False)), ('generic', dict(generic_kernel=True))]


def _case(case):
    name, d, N, T, tkind, prior, kw = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    tgt = make_target(dict(kind=tkind, d=d))
    prior = dict(prior, d=d)
    pk, pa, pb = prior_arrays(prior, d)
    nseed = 2 * kw.get('DEpairs', 1) * N + 7
    if prior['kind'] == 'uniform':
        hist = np.array(prior['loc']) + np.array(prior['scale']) * rng.uniform(size=(nseed, d))
        if not kw.get('hardboundaries', True):
            hist = hist + .8 * np.array(prior['scale']) * rng.normal(size=(nseed, d))   # wide archive: most jumps leave the support
            hist[:N] = np.array(prior['loc']) + np.array(prior['scale']) * rng.uniform(size=(N, d))   # starts inside
    elif prior['kind'] == 'norm':
        hist = np.array(prior['loc']) + np.array(prior['scale']) * rng.normal(size=(nseed, d))
    else:
        hist = rng.normal(size=(nseed, d)) * 2.0
    return name, d, N, T, tgt, (pk, pa, pb), kw, hist


@pytest.mark.parametrize('form', FORMS, ids=[f[0] for f in FORMS])
@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_multitry_matches_c_oracle(case, form):
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    name, d, N, T, tgt, (pk, pa, pb), kw, hist = _case(case)
    starts = hist[:N].copy()
    okw = dict(kw, adapt_crossover=False)
    ref = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=77, prior_kind=pk, prior_a=pa, prior_b=pb,
                                 **okw).run(T, rows_dbg_n=96)
    eng = DreamEngine(d, N, hist, starts, tgt, pk, pa, pb, seed=77, **okw, **form[1])
    if form[0] == 'two_stage':
        assert eng.draw_ws is not None
    trace, logp, dec = eng.run(T)
    got_dec = dec.t().contiguous().cpu().numpy().astype(np.uint32)
    got_lp = logp.t().contiguous().cpu().numpy()
    got_sp = trace[:, :, :d].permute(1, 0, 2).contiguous().cpu().numpy()
    assert np.array_equal(got_dec, ref['decisions']), 'decisions differ at %s' % (np.argwhere(got_dec != ref['decisions'])[:3].tolist(),)
    fin = np.isfinite(ref['logp'])
    assert np.array_equal(fin, np.isfinite(got_lp))
    err = np.abs(got_lp[fin] - ref['logp'][fin])
    assert np.all(err <= 10 * logp_tol(ref['logp'][fin])), err.max()
    np.testing.assert_allclose(got_sp, ref['states'], rtol=1e-10, atol=1e-11)
    if 'regen' in name:      # the case is only worth its name if batches were in fact regenerated: more archive rows
        k, snk = kw['multitry'], (ref['decisions'] >> 1) & 1          # sampled than the 2k-1 points of an iteration need
        normal = np.where(snk == 1, 3 * (2 * k - 1), 2 * kw.get('DEpairs', 1) * (2 * k - 1))
        nreg = int(((ref['rows'] >= 0).sum(axis=2) > normal).sum())
        assert nreg > 0, 'no regenerated batch in this case'


def test_two_stage_runs_long_windows():
    """history_thin larger than one launch of the fused kernel used to be: the scratch covers the whole window."""
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    d, N, T = 10, 128, 70
    rng = np.random.default_rng(5)
    tgt = make_target(dict(kind='mixture', d=d))
    hist = rng.normal(size=(2 * N + 3, d))
    kw = dict(multitry=5, snooker=.1, history_thin=33, adapt_crossover=False)
    ref = c_oracle.OracleSampler(d, N, hist, hist[:N].copy(), tgt.kind, tgt.table(), seed=3, **kw).run(T)
    eng = DreamEngine(d, N, hist, hist[:N].copy(), tgt, seed=3, **kw)
    trace, logp, dec = eng.run(T)
    assert np.array_equal(dec.t().contiguous().cpu().numpy().astype(np.uint32), ref['decisions'])


# ---------------------------------------------------------------- two-stage single-try step (dreamzs_st2_kernel.cuh)
ST_CASES = [
    ('banana200', 200, 64, 24, 'banana', dict(kind='flat'), dict(snooker=.1, history_thin=6)),
    ('banana530_r8', 530, 20, 10, 'banana', dict(kind='flat'), dict(snooker=.3, history_thin=2)),
    ('mix10_pairs3', 10, 60, 30, 'mixture', dict(kind='flat'), dict(snooker=.2, history_thin=5, DEpairs=3, gamma_levels=2)),
    ('sum6_bounds', 6, 80, 40, 'sumshift', dict(kind='uniform', loc=[-.5] * 6, scale=[1.] * 6), dict(snooker=.2, history_thin=5, zeta=1e-3, lamb=.4)),
    ('norm5', 5, 40, 30, 'sumshift', dict(kind='norm', loc=[.3] * 5, scale=[2.] * 5), dict(snooker=.1, history_thin=4)),
    ('gauss40_norm', 40, 40, 20, 'gaussian', dict(kind='norm', loc=[5.] * 40, scale=[10.] * 40), dict(snooker=.1, history_thin=5)),
]


@pytest.mark.parametrize('draw_iters', [None, 2])
@pytest.mark.parametrize('case', ST_CASES, ids=[c[0] for c in ST_CASES])
def test_two_stage_single_try_matches_c_oracle(case, draw_iters):
    """draw kernel + chain kernel per sub-span of a window (draw_iters=2: windows cut into sub-spans of two iterations)."""
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    name, d, N, T, tgt, (pk, pa, pb), kw, hist = _case(tuple(case[:6]) + (dict(case[6], multitry=1),))
    kw = dict(kw, adapt_crossover=False)
    ref = c_oracle.OracleSampler(d, N, hist, hist[:N].copy(), tgt.kind, tgt.table(), seed=78, prior_kind=pk, prior_a=pa,
                                 prior_b=pb, **kw).run(T)
    eng = DreamEngine(d, N, hist, hist[:N].copy(), tgt, pk, pa, pb, seed=78, draw_iters=draw_iters, **kw)
    assert eng.draw_ws is not None
    launches0 = eng.launches
    trace, logp, dec = eng.run(T)
    assert eng.launches - launches0 > T // kw['history_thin']          # more than one kernel per window: the two-stage form ran
    got_dec = dec.t().contiguous().cpu().numpy().astype(np.uint32)
    got_lp = logp.t().contiguous().cpu().numpy()
    got_sp = trace[:, :, :d].permute(1, 0, 2).contiguous().cpu().numpy()
    assert np.array_equal(got_dec, ref['decisions']), 'decisions differ at %s' % (np.argwhere(got_dec != ref['decisions'])[:3].tolist(),)
    assert np.all(np.abs(got_lp - ref['logp']) <= 10 * logp_tol(ref['logp'])), (np.abs(got_lp - ref['logp']) / logp_tol(ref['logp'])).max()
    np.testing.assert_allclose(got_sp, ref['states'], rtol=1e-10, atol=1e-11)
