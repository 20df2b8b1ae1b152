"""Writes the golden vectors in this directory by running the UNMODIFIED reference
(/root/reference/pydream) in lock-step with injected counter-based RNG (oracle/ref_harness.py).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Each ``<case>.npz`` holds the inputs (seed archive, starts, JSON options) and the reference's outputs
(states, log_prior, log_like, accept flags, archive rows drawn, multinomial picks, final archive,
adapted probabilities).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H          # noqa: E402
from pydream_b200 import targets as T        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


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


def make_params(prior):
    def f(P):
        from scipy.stats import norm, uniform
        if prior['kind'] == 'flat':
            return [P.FlatParam(test_value=np.zeros(prior['d']))]
        if prior['kind'] == 'norm':
            return [P.SampledParam(norm, loc=np.array(prior['loc']), scale=np.array(prior['scale']))]
        if prior['kind'] == 'uniform':
            return [P.SampledParam(uniform, loc=np.array(prior['loc']), scale=np.array(prior['scale']))]
        if prior['kind'] == 'mixed':   # two parameter groups: norm then uniform
            return [P.SampledParam(norm, loc=np.array(prior['loc'][0]), scale=np.array(prior['scale'][0])),
                    P.SampledParam(uniform, loc=np.array(prior['loc'][1]), scale=np.array(prior['scale'][1]))]
        raise ValueError(prior)
    return f


CASES = {
    # C2-like: dense Gaussian, DE only
    'gauss100_de': dict(target=dict(kind='gaussian', d=100), prior=dict(kind='flat', d=100), N=8, T=40, nseed=40,
                        seed=11, hist='lhs', kw=dict(snooker=0, history_thin=5)),
    'gauss30_snooker': dict(target=dict(kind='gaussian', d=30), prior=dict(kind='flat', d=30), N=6, T=60, nseed=30,
                            seed=12, hist='lhs', kw=dict(snooker=.3, history_thin=10)),
    # C3-like: mixture, multi-try 5 + snooker
    'mix10_mt5': dict(target=dict(kind='mixture', d=10), prior=dict(kind='flat', d=10), N=6, T=60, nseed=24,
                      seed=13, hist='normal', kw=dict(multitry=5, snooker=.1, history_thin=3)),
    'gauss20_mt3_snk': dict(target=dict(kind='gaussian', d=20), prior=dict(kind='flat', d=20), N=5, T=50, nseed=20,
                            seed=14, hist='lhs', kw=dict(multitry=3, snooker=.5, history_thin=1)),
    # C4-like: banana, several DE pairs
    'banana8_de3': dict(target=dict(kind='banana', d=8), prior=dict(kind='flat', d=8), N=7, T=60, nseed=60,
                        seed=15, hist='banana', kw=dict(DEpairs=3, snooker=.1, history_thin=4)),
    # C5-like: crossover adaptation during burn-in (normal prior, reference test model)
    'norm4_adaptcr': dict(target=dict(kind='sumshift', d=4),
                          prior=dict(kind='norm', loc=[-6.6, 3, 1.0, -.12], scale=[.13, 5, .9, 1.0]), N=5, T=90,
                          nseed=20, seed=16, hist='prior_norm', kw=dict(adapt_crossover=True, crossover_burnin=50,
                                                                       history_thin=2, nCR=3)),
    'gauss12_adapt_gamma': dict(target=dict(kind='gaussian', d=12), prior=dict(kind='flat', d=12), N=6, T=80,
                                nseed=24, seed=17, hist='lhs',
                                kw=dict(adapt_crossover=True, adapt_gamma=True, gamma_levels=4, crossover_burnin=45,
                                        nCR=4, history_thin=5, snooker=.2)),
    # hard boundaries (reference test model multidmodel_uniform) with large jumps
    'unif4_bounds': dict(target=dict(kind='sumshift', d=4),
                         prior=dict(kind='uniform', loc=[-5, -9, 5, 3], scale=[15, 11, 2, 5]), N=5, T=120, nseed=20,
                         seed=18, hist='prior_unif', kw=dict(history_thin=2, p_gamma_unity=.5, lamb=.5)),
    'unif4_bounds_mt3': dict(target=dict(kind='sumshift', d=4),
                             prior=dict(kind='uniform', loc=[-5, -9, 5, 3], scale=[15, 11, 2, 5]), N=5, T=60,
                             nseed=20, seed=19, hist='prior_unif', kw=dict(history_thin=2, p_gamma_unity=.5,
                                                                         multitry=3, snooker=.2)),
    # C1: 10-D standard normal prior, likelihood == 0, defaults
    'stdnorm10_defaults': dict(target=dict(kind='constant', d=10),
                               prior=dict(kind='norm', loc=[0.] * 10, scale=[1.] * 10), N=3, T=150, nseed=100,
                               seed=20, hist='prior_norm', kw=dict()),
    # one-dimensional models (pydream/tests/test_dream.py onedmodel: nCR collapses to 1, no crossover adaptation,
    # scalar boundary mask Dream.py:84-85)
    'norm1_onedim': dict(target=dict(kind='sumshift', d=1), prior=dict(kind='norm', loc=[-2.], scale=[3.]), N=4, T=60,
                         nseed=16, seed=41, hist='prior_norm',
                         kw=dict(history_thin=2, crossover_burnin=20, nCR=1, adapt_crossover=False, snooker=.2)),
    # (with one dimension the reference itself crashes under numpy 2 when multi-try is on -- np.squeeze leaves a 0-d
    # q_proposal, Dream.py:723 -- and when a proposal leaves a finite boundary -- Dream.py:766: no golden case possible)
    'mixed5_nobounds': dict(target=dict(kind='constant', d=5),
                            prior=dict(kind='mixed', loc=[[0., 1.], [-1., -2., 0.]], scale=[[1., 2.], [2., 4., 1.]]),
                            N=4, T=80, nseed=16, seed=21, hist='mixed',
                            kw=dict(hardboundaries=False, multitry=3, history_thin=2, lamb=.3, p_gamma_unity=.6)),
}


# parallel tempering (pydream/core.py:131-236): the reference's own driver on the lock-step pool
PT_CASES = {
    'pt_gauss8_snooker': dict(target=dict(kind='gaussian', d=8), prior=dict(kind='flat', d=8), N=6, T=42, nseed=30,
                              seed=31, hist='lhs', kw=dict(snooker=.2, history_thin=3, adapt_crossover=False)),
    'pt_norm4_adaptcr_mt3': dict(target=dict(kind='sumshift', d=4),
                                 prior=dict(kind='norm', loc=[-6.6, 3, 1.0, -.12], scale=[.13, 5, .9, 1.0]), N=5, T=60,
                                 nseed=20, seed=32, hist='prior_norm',
                                 kw=dict(adapt_crossover=True, crossover_burnin=30, history_thin=2, nCR=3, multitry=3,
                                         snooker=.1)),
    'pt_mix10_de2': dict(target=dict(kind='mixture', d=10), prior=dict(kind='flat', d=10), N=8, T=40, nseed=40,
                         seed=33, hist='normal', kw=dict(DEpairs=2, snooker=0, history_thin=4, adapt_crossover=False)),
}


def make_history(kind, nseed, d, prior, rng):
    if kind == 'lhs':
        return rng.uniform(-5, 15, size=(nseed, d))
    if kind == 'normal':
        return rng.normal(size=(nseed, d))
    if kind == 'banana':
        h = rng.normal(size=(nseed, d))
        h[:, 0] *= 10
        return h
    if kind == 'prior_norm':
        return np.array(prior['loc']) + np.array(prior['scale']) * rng.normal(size=(nseed, d))
    if kind == 'prior_unif':
        return np.array(prior['loc']) + np.array(prior['scale']) * rng.uniform(size=(nseed, d))
    if kind == 'mixed':
        loc = np.concatenate([np.array(x) for x in prior['loc']])
        scale = np.concatenate([np.array(x) for x in prior['scale']])
        h = loc + scale * rng.uniform(size=(nseed, d))
        return h
    raise ValueError(kind)


def main(only=None):
    for name, c in CASES.items():
        if only and name not in only:
            continue
        d = c['target']['d']
        rng = np.random.default_rng(c['seed'])
        hist = make_history(c['hist'], c['nseed'], d, c['prior'], rng)
        starts = hist[:c['N']].copy()
        tgt = make_target(c['target'])
        out = H.run_lockstep(make_params(c['prior']), tgt, c['N'], c['T'], starts, hist, seed=c['seed'], **c['kw'])
        meta = dict(target=c['target'], prior=c['prior'], N=c['N'], T=c['T'], seed=c['seed'], kw=c['kw'])
        np.savez_compressed(os.path.join(HERE, name + '.npz'), meta=json.dumps(meta), history=hist, starts=starts, **out)
        print(name, 'accept rate %.3f' % out['accept'].mean(), 'rows', int(out['count_final']), flush=True)
    for name, c in PT_CASES.items():
        if only and name not in only:
            continue
        d = c['target']['d']
        rng = np.random.default_rng(c['seed'])
        hist = make_history(c['hist'], c['nseed'], d, c['prior'], rng)
        starts = hist[:c['N']].copy()
        tgt = make_target(c['target'])
        out = H.run_lockstep_pt(make_params(c['prior']), tgt, c['N'], c['T'], starts, hist, seed=c['seed'], **c['kw'])
        meta = dict(target=c['target'], prior=c['prior'], N=c['N'], T=c['T'], seed=c['seed'], kw=c['kw'], tempering=True)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), meta=json.dumps(meta), history=hist, starts=starts, **out)
        sp = out['sampled_params']
        print(name, 'swaps accepted %d of %d' % (int(np.any(sp[:, 0::2] != sp[:, 1::2], axis=(0, 2)).sum()), c['T']),
              'rows', int(out['count_final']), flush=True)


if __name__ == '__main__':
    main(sys.argv[1:])
