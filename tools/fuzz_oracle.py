"""Differential fuzzing of the C restatement (oracle/dreamzs_oracle.c) against the UNMODIFIED reference under the
lock-step harness, over random option combinations wider than the committed golden cases (adaptation of crossover and
gamma levels, priors of every closed form, no hard boundaries, chains starting ON archive rows so that snooker
projections hit D = 0, parallel tempering).  Build container only (needs /root/reference).
    python tools/fuzz_oracle.py [ncases] [first_seed]"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

import make_golden as G                                                     # noqa: E402
from golden_util import make_target, prior_arrays, sampler_kwargs, decode_decisions, logp_tol   # noqa: E402
from oracle import c_oracle, ref_harness as H                               # noqa: E402


def random_case(i):
    rng = np.random.default_rng(i)
    kind = ['gaussian', 'banana', 'mixture', 'sumshift', 'constant'][int(rng.integers(0, 5))]
    d = int(rng.integers(2, 15))
    DEpairs = int(rng.integers(1, 4))
    N = int(rng.integers(2 * DEpairs + 1, 2 * DEpairs + 6))
    thin = int(rng.integers(1, 6))
    T = thin * int(rng.integers(5, 14))
    burn = int(rng.integers(12, max(13, T - 2))) if T > 16 else 3
    kw = dict(snooker=float(rng.choice([0., .1, .3, .7])), history_thin=thin, DEpairs=DEpairs,
              multitry=[False, False, 3, 5, 4][int(rng.integers(0, 5))], nCR=int(rng.integers(1, min(d, 5) + 1)),
              adapt_crossover=bool(rng.integers(0, 2)), adapt_gamma=bool(rng.integers(0, 2)), crossover_burnin=burn,
              p_gamma_unity=float(rng.choice([0., .2, .5, 1.])), gamma_levels=int(rng.integers(1, 5)),
              lamb=float(rng.choice([.05, .4])), zeta=float(rng.choice([1e-12, 1e-3])),
              hardboundaries=bool(rng.integers(0, 4) > 0))
    pk = ['flat', 'norm', 'uniform', 'mixed'][int(rng.integers(0, 4))] if kind in ('sumshift', 'constant') else 'flat'
    if pk == 'mixed' and d < 4:      # the reference cannot stack the bounds of a one-element array parameter (Dream.py:104)
        pk = 'uniform'
    nseed = 2 * DEpairs * N + int(rng.integers(0, 12))
    if pk == 'flat':
        prior = dict(kind='flat', d=d)
        hist = rng.normal(size=(nseed, d)) * float(rng.choice([1., 4.]))
    elif pk == 'norm':
        prior = dict(kind='norm', loc=list(rng.normal(size=d)), scale=list(rng.uniform(.2, 3., size=d)))
        hist = np.array(prior['loc']) + np.array(prior['scale']) * rng.normal(size=(nseed, d))
    elif pk == 'uniform':
        prior = dict(kind='uniform', loc=list(rng.normal(size=d)), scale=list(rng.uniform(.5, 6., size=d)))
        hist = np.array(prior['loc']) + np.array(prior['scale']) * rng.uniform(size=(nseed, d))
    else:
        n0 = int(rng.integers(2, d - 1))
        prior = dict(kind='mixed', loc=[list(rng.normal(size=n0)), list(rng.normal(size=d - n0))],
                     scale=[list(rng.uniform(.2, 3., size=n0)), list(rng.uniform(.5, 6., size=d - n0))])
        loc = np.concatenate(prior['loc']); sc = np.concatenate(prior['scale'])
        hist = loc + sc * rng.uniform(size=(nseed, d))
    return dict(target=dict(kind=kind, d=d), prior=prior, N=N, T=T, seed=9000 + i, kw=kw, tempering=bool(rng.integers(0, 4) == 0)), hist


def check(meta, hist):
    d, N, T = meta['target']['d'], meta['N'], meta['T']
    starts = hist[:N].copy()
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    s = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=meta['seed'], prior_kind=pk, prior_a=pa,
                               prior_b=pb, **sampler_kwargs(meta))
    if meta['tempering']:
        z = H.run_lockstep_pt(G.make_params(meta['prior']), tgt, N, T, starts, hist, seed=meta['seed'], **meta['kw'])
        out = s.run_pt(T)
        assert np.array_equal(out['swaps'][:, :2], z['pairs']), 'swap pairs'
        ref_sp, got_sp = z['sampled_params'], out['sampled_params']
        assert np.array_equal(got_sp != got_sp[:, [0] + list(range(0, 2 * T - 1))], ref_sp != ref_sp[:, [0] + list(range(0, 2 * T - 1))]), 'change pattern'
        np.testing.assert_allclose(got_sp, ref_sp, rtol=1e-9, atol=1e-10)
        ref_lp = z['log_ps'][:, :, 0]
        assert np.all(np.abs(out['log_ps'] - ref_lp) <= 10 * logp_tol(ref_lp)), 'log_ps'
    else:
        z = H.run_lockstep(G.make_params(meta['prior']), tgt, N, T, starts, hist, seed=meta['seed'], **meta['kw'])
        out = s.run(T)
        dec = decode_decisions(out['decisions'])
        assert np.array_equal(dec['changed'], z['accept']), 'accept sequence'
        ref_lp = z['log_like'] + z['log_prior']
        assert np.all(np.abs(out['logp'] - ref_lp) <= 10 * logp_tol(ref_lp)), 'logp'
        np.testing.assert_allclose(out['states'], z['states'], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(s.gamma_probs, z['gamma_probs'][-1], rtol=1e-10)
    np.testing.assert_allclose(s.history_flat, z['history_final'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(s.cr_probs, z['cr_probs'][-1], rtol=1e-10)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for i in range(first, first + n):
        meta, hist = random_case(i)
        try:
            check(meta, hist)
        except Exception as e:      # noqa: BLE001
            bad += 1
            msg = traceback.format_exc().strip().splitlines()
            print('case %d FAILED: %s\n   %s\n   %s' % (i, meta, msg[-1][:300], ' | '.join(l.strip() for l in msg[-6:-1])[:600]), flush=True)
    print('%d cases, %d failed' % (n, bad))


if __name__ == '__main__':
    main()
