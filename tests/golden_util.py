"""Helpers shared by the golden-vector tests (CPU oracle and CUDA path)."""
import glob
import json
import os

import numpy as np

from pydream_b200 import targets as T

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _all_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def golden_cases():
    return [c for c in _all_cases() if not c.startswith('pt_')]


def golden_pt_cases():
    """Parallel-tempering cases (the reference's _sample_dream_pt driver, pydream/core.py:131-236)."""
    return [c for c in _all_cases() if c.startswith('pt_')]


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(str(z['meta']))
    return meta, z


def make_target(spec):
    kind = spec['kind']
    if kind == 'gaussian':
        return T.CorrelatedGaussian.benchmark(spec['d'])
    if kind == 'mixture':
        return T.BimodalMixture.benchmark(spec['d'])
    if kind == 'banana':
        return T.Banana(spec['d'], spec.get('b', 0.1))
    if kind == 'sumshift':
        return T.SumShift(spec['d'], 3.0)
    if kind == 'constant':
        return T.Constant(spec['d'], 0.0)
    raise ValueError(kind)


def prior_arrays(prior, d):
    """-> (kind[d] int32, a[d], b[d]) in the encoding of include/dreamzs.h."""
    kind = np.zeros(d, dtype=np.int32)
    a, b = np.zeros(d), np.ones(d)
    if prior['kind'] == 'norm':
        kind[:] = 1
        a[:], b[:] = prior['loc'], prior['scale']
    elif prior['kind'] == 'uniform':
        kind[:] = 2
        a[:], b[:] = prior['loc'], prior['scale']
    elif prior['kind'] == 'mixed':
        n0 = len(prior['loc'][0])
        kind[:n0], kind[n0:] = 1, 2
        a[:] = np.concatenate(prior['loc'])
        b[:] = np.concatenate(prior['scale'])
    return kind, a, b


def sampler_kwargs(meta):
    """Dream kwargs of a golden case -> kwargs of OracleSampler / the engine."""
    kw = dict(meta['kw'])
    T_ = meta['T']
    out = dict(nCR=kw.get('nCR', 3), gamma_levels=kw.get('gamma_levels', 1), DEpairs=kw.get('DEpairs', 1),
               multitry=kw.get('multitry', 1) or 1, snooker=kw.get('snooker', .1),
               p_gamma_unity=kw.get('p_gamma_unity', .2), lamb=kw.get('lamb', .05), zeta=kw.get('zeta', 1e-12),
               history_thin=kw.get('history_thin', 10), hardboundaries=kw.get('hardboundaries', True),
               adapt_crossover=kw.get('adapt_crossover', True), adapt_gamma=kw.get('adapt_gamma', False),
               crossover_burnin=kw.get('crossover_burnin', T_ // 10))
    if out['multitry'] is True:
        out['multitry'] = 5
    return out


def decode_decisions(dec):
    dec = np.asarray(dec)
    return dict(changed=dec & 1, snooker=(dec >> 1) & 1, cr=(dec >> 2) & 15, lvl=(dec >> 6) & 15, delta=(dec >> 10) & 15,
                sel=(dec >> 14) & 15, gamma_one=(dec >> 18) & 1, accepted=(dec >> 19) & 1)


def logp_tol(ref):
    """north_star tolerance: 1e-12, relative to max(1, |logp|) (ulp(1e4) alone is 1.8e-12)."""
    return 1e-12 * np.maximum(1.0, np.abs(ref))
