"""Parallel tempering on the GPU: dreamzs_step_tempered + dreamzs_pt_swap through the engine
and run_dream(..., tempering=True), against (a) the golden vectors written by the reference's own driver
_sample_dream_pt (pydream/core.py:131-236) on the lock-step pool and (b) the C oracle on larger seeded inputs.
Swap pairs, swap verdicts and accept/reject sequences bit-exact; log_ps within 1e-12 * max(1, |logp|)."""
import numpy as np
import pytest

from golden_util import golden_pt_cases, load_case, make_target, prior_arrays, sampler_kwargs, logp_tol

pytestmark = pytest.mark.gpu


def _tempered(eng, niter, temperature=None):
    trace, logp, dec, swaps = eng.run_tempered(niter, temperature)
    return (trace[:, :, :eng.d].contiguous().cpu().numpy(), logp.cpu().numpy(), dec.cpu().numpy().astype(np.uint32),
            swaps.cpu().numpy())


@pytest.mark.parametrize('name', golden_pt_cases())
def test_cuda_tempering_matches_reference_golden(name):
    from pydream_b200.engine import DreamEngine, temperature_ladder
    meta, z = load_case(name)
    d = meta['target']['d']
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    eng = DreamEngine(d, meta['N'], z['history'], z['starts'], tgt, pk, pa, pb, seed=meta['seed'], **sampler_kwargs(meta))
    np.testing.assert_array_equal(temperature_ladder(meta['N']), z['T'])
    sp, lp, dec, swaps = _tempered(eng, meta['T'])
    ref_sp, ref_lp = z['sampled_params'], z['log_ps'][:, :, 0]
    assert sp.shape == ref_sp.shape                                                   # (N, 2 niter, d), core.py:145
    np.testing.assert_array_equal(swaps[:, :2].astype(np.int64), z['pairs'])
    ref_swapped = np.any(ref_sp[:, 0::2] != ref_sp[:, 1::2], axis=(0, 2))
    np.testing.assert_array_equal(np.any(sp[:, 0::2] != sp[:, 1::2], axis=(0, 2)), ref_swapped)
    ref_acc = np.any(ref_sp[:, 2::2] != ref_sp[:, 1:-1:2], axis=2)
    np.testing.assert_array_equal((dec[:, 2::2] & 1).astype(bool), ref_acc)
    np.testing.assert_allclose(sp, ref_sp, rtol=1e-10, atol=1e-11)
    assert np.all(np.abs(lp - ref_lp) <= logp_tol(ref_lp)), np.abs(lp - ref_lp).max()
    hf = eng.history_flat()
    assert hf.shape == z['history_final'].shape
    np.testing.assert_allclose(hf, z['history_final'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.cr_probs.cpu().numpy(), z['cr_probs'][-1], rtol=1e-10)


CASES = [
    ('pt_gauss100', 100, 64, 40, 'gaussian', dict(snooker=.1, history_thin=10)),          # dense Gaussian on the generic kernel
    ('pt_mix10_mt5', 10, 96, 30, 'mixture', dict(multitry=5, snooker=.1, history_thin=5)),
    ('pt_banana200', 200, 24, 20, 'banana', dict(snooker=.2, history_thin=4)),
    ('pt_gauss50_adapt', 50, 48, 48, 'gaussian', dict(snooker=.1, history_thin=6, adapt_crossover=True, crossover_burnin=30)),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_cuda_tempering_matches_c_oracle(case):
    import zlib
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    from pydream_b200 import _cabi
    name, d, N, T, tkind, kw = case
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    tgt = make_target(dict(kind=tkind, d=d))
    nseed = 2 * N + 11
    hist = rng.uniform(-5, 15, size=(nseed, d)) if tkind != 'mixture' else rng.normal(size=(nseed, d))
    starts = hist[:N].copy()
    okw = dict(adapt_crossover=False, adapt_gamma=False, crossover_burnin=0)
    okw.update(kw)
    orc = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=9, nthreads=4, **okw)
    ref = orc.run_pt(T)
    eng = DreamEngine(d, N, hist, starts, tgt, seed=9, **okw)
    sp, lp, dec, swaps = _tempered(eng, T)
    np.testing.assert_array_equal(swaps[:, :3].astype(np.int64), ref['swaps'])           # pair and verdict
    np.testing.assert_array_equal(dec, ref['decisions'])
    assert ((dec[:, 1::2] & _cabi.DECISION_SWAPPED) != 0).sum() == 2 * ref['swaps'][:, 2].sum()
    err = np.abs(lp - ref['log_ps'])
    assert np.all(err <= logp_tol(ref['log_ps'])), (err / logp_tol(ref['log_ps'])).max()
    np.testing.assert_allclose(sp, ref['sampled_params'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.history_flat(), orc.history_flat, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.cr_probs.cpu().numpy(), orc.cr_probs, rtol=1e-10)


def test_unit_temperatures_reproduce_the_untempered_step():
    """With T = 1 for every chain the tempered step consumes the same draws and takes the same decisions as the
    fused step (here: the dense-Gaussian window kernel).  Compared on the first iteration: at equal temperatures
    alpha = 0, so every proposed exchange is accepted (log u < 0) and the chains are permuted from then on."""
    from pydream_b200.engine import DreamEngine
    from pydream_b200 import targets
    d, N, T = 100, 32, 6
    rng = np.random.default_rng(4)
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(3 * N, d))
    kw = dict(seed=2, snooker=.1, history_thin=5)
    a = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    tr_a, lp_a, dec_a = a.run(1)
    b = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    sp, lp, dec, swaps = _tempered(b, T, np.ones(N))
    assert np.all(swaps[:, 2] == 1) and np.all(swaps[:, 3] == 0)
    np.testing.assert_array_equal(dec[:, 0], dec_a.cpu().numpy().astype(np.uint32)[:, 0])
    np.testing.assert_allclose(sp[:, 0], tr_a[:, 0, :d].cpu().numpy(), rtol=1e-10, atol=1e-11)
    ref_lp = lp_a.cpu().numpy()[:, 0]
    assert np.all(np.abs(lp[:, 0] - ref_lp) <= logp_tol(ref_lp))
    # every record after an exchange is a permutation of the record before it
    for t in range(T):
        i, j = int(swaps[t, 0]), int(swaps[t, 1])
        np.testing.assert_array_equal(sp[i, 2 * t + 1], sp[j, 2 * t])
        np.testing.assert_array_equal(sp[j, 2 * t + 1], sp[i, 2 * t])
        others = [c for c in range(N) if c not in (i, j)]
        np.testing.assert_array_equal(sp[others, 2 * t + 1], sp[others, 2 * t])


def test_run_dream_tempering_shapes_and_swaps(tmp_path, monkeypatch):
    """run_dream(..., tempering=True) returns arrays (nchains, 2 niterations, ndim) / (nchains, 2 niterations, 1)
    (pydream/core.py:145-146, 236); the post-swap record of an iteration is a permutation of its post-step record."""
    from scipy.stats import norm
    from pydream_b200 import targets
    from pydream_b200.core import run_dream
    from pydream_b200.parameters import SampledParam
    monkeypatch.chdir(tmp_path)
    mu, sd = np.array([-6.6, 3, 1.0, -.12]), np.array([.13, 5, .9, 1.0])
    nchains, niter = 6, 40
    sampled, log_ps = run_dream([SampledParam(norm, loc=mu, scale=sd)], targets.SumShift(4, 3.0), nchains=nchains,
                                niterations=niter, tempering=True, history_thin=2, model_name='pt', verbose=False, seed=8)
    assert isinstance(sampled, np.ndarray) and sampled.shape == (nchains, 2 * niter, 4) and log_ps.shape == (nchains, 2 * niter, 1)
    nswaps = 0
    for t in range(niter):
        step, swap = sampled[:, 2 * t], sampled[:, 2 * t + 1]
        moved = np.where(np.any(step != swap, axis=1))[0]
        assert len(moved) in (0, 2)
        if len(moved) == 2:
            nswaps += 1
            a, b = moved
            np.testing.assert_array_equal(swap[a], step[b])
            np.testing.assert_array_equal(swap[b], step[a])
            np.testing.assert_array_equal(log_ps[a, 2 * t + 1], log_ps[b, 2 * t])
    assert nswaps > 0
    history = np.load('pt_DREAM_chain_history.npy')
    assert len(history) == 4 * (nchains * niter // 2 + 40)


def test_astep_operator_contract():
    """DreamEngine.astep(q0, T, last_loglike, last_logprior) -> (q_new, log_prior, log_like), the reference's step
    operator (Dream.py:193, 422) for all chains at once: a loop over astep reproduces run(); passing the returned
    state back in (as _sample_dream_pt_chain does, core.py:239-246) changes nothing; logp monotone on acceptance at
    T = 1 without snooker (test_astep_*, pydream/tests/test_dream.py:499-562)."""
    import torch
    from pydream_b200.engine import DreamEngine
    from pydream_b200 import targets
    d, N, T = 100, 48, 14
    rng = np.random.default_rng(6)
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(3 * N, d))
    kw = dict(seed=3, snooker=0., history_thin=5)
    a = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    tr_a, lp_a, dec_a = a.run(T)
    b = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    c = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    q = lk = pr = None
    for t in range(T):
        q_b, pr_b, lk_b = b.astep()                                     # continue from the engine's own state
        q, pr, lk = c.astep(q, 1., lk, pr) if q is not None else c.astep()   # state handed back in every call
        assert q_b.shape == (N, d) and pr_b.shape == (N,) and lk_b.shape == (N,)
        ref_lp = lp_a[:, t]
        for qq, pp, ll in ((q_b, pr_b, lk_b), (q, pr, lk)):
            np.testing.assert_allclose(qq.cpu().numpy(), tr_a[:, t, :d].cpu().numpy(), rtol=1e-10, atol=1e-11)
            assert np.all(np.abs((ll + pp - ref_lp).cpu().numpy()) <= logp_tol(ref_lp.cpu().numpy()))
        np.testing.assert_array_equal(b.last_decisions.cpu().numpy(), dec_a[:, t].cpu().numpy())
        if t > 0:
            acc = (dec_a[:, t] & 1).bool()
            # accepted points were tested with log u < logp_new - logp_old; rejected ones keep their logp
            assert torch.all(lp_a[:, t][~acc] == lp_a[:, t - 1][~acc])
    assert b.archive_rows == a.archive_rows and b.iter == a.iter == T
    # a tempered step through the same operator: T -> 0 flattens the likelihood, so (flat prior) every finite proposal is accepted
    e = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    q_e, pr_e, lk_e = e.astep(T=1e-300)
    assert int((e.last_decisions & 1).sum()) >= N - 2
