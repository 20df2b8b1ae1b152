"""BASELINE.json's full single-GPU sizes: (a) the CUDA path against the C oracle on the same seeded inputs
(decision words bit-exact, log-posteriors within the north_star tolerance, states rtol 1e-10) -- C2 1024 x 100,
C3 4096 x 10 multi-try 5, C4's per-GPU share 4096 x 200, C5-shaped 8192 x 50 with crossover adaptation; the oracle
runs on the host threads and takes seconds at these sizes -- and (b) size-independent properties: the archive tail
is exactly the states at the appending iterations, rejected steps repeat the previous state, the log-posterior
trace is the target evaluated on the trace, and two independent kernels agree."""
import numpy as np
import pytest

from golden_util import logp_tol, decode_decisions
from pydream_b200 import targets

pytestmark = pytest.mark.gpu


def _oracle_parity(d, N, T, tgt, hist, kw, seed):
    """The same run through the CUDA path (C ABI) and the C oracle; returns the oracle's result dict."""
    import os
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    starts = hist[:N].copy()
    nthreads = max(1, min(32, len(os.sched_getaffinity(0))))
    orc = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=seed, nthreads=nthreads, **kw)
    ref = orc.run(T)
    eng = DreamEngine(d, N, hist, starts, tgt, seed=seed, **kw)
    trace, logp, dec = eng.run(T)
    got_dec = dec.t().contiguous().cpu().numpy().astype(np.uint32)
    got_logp = logp.t().contiguous().cpu().numpy()
    got_states = trace[:, :, :d].permute(1, 0, 2).contiguous().cpu().numpy()
    np.testing.assert_array_equal(got_dec, ref['decisions'])                     # accept / reject, snooker, CR, gamma level, pick
    err = np.abs(got_logp - ref['logp'])
    assert np.all(err <= logp_tol(ref['logp'])), (err / logp_tol(ref['logp'])).max()
    np.testing.assert_allclose(got_states, ref['states'], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.history_flat(), orc.history_flat, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(eng.cr_probs.cpu().numpy(), orc.cr_probs, rtol=1e-10)
    return ref


def test_c2_full_size_matches_oracle():
    """BASELINE config 2: 100-D correlated Gaussian, 1024 chains, snooker .1, thin 10, 200 iterations (20 windows of
    the window kernel, 5 refreshes of the carried state), archive seed 2N rows."""
    d, N, T = 100, 1024, 200
    rng = np.random.default_rng(142)
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(2 * N, d))
    ref = _oracle_parity(d, N, T, tgt, hist, dict(snooker=.1, history_thin=10), seed=21)
    assert 0.02 < (ref['decisions'] & 1).mean() < 0.9


def test_c3_full_size_matches_oracle():
    """BASELINE config 3: 10-D bimodal mixture, 4096 chains, multi-try 5 + snooker, 40 iterations."""
    d, N, T = 10, 4096, 40
    rng = np.random.default_rng(143)
    tgt = targets.BimodalMixture.benchmark(d)
    hist = rng.normal(size=(2 * N, d))
    _oracle_parity(d, N, T, tgt, hist, dict(snooker=.1, history_thin=10, multitry=5), seed=22)


def test_c4_share_matches_oracle():
    """BASELINE config 4, the per-GPU share at 2 GPUs: 200-D twisted Gaussian, 4096 chains, 30 iterations."""
    d, N, T = 200, 4096, 30
    rng = np.random.default_rng(144)
    tgt = targets.Banana(d, 0.1)
    hist = rng.normal(size=(2 * N, d)) * np.sqrt(np.concatenate([[100.0], np.ones(d - 1)]))
    _oracle_parity(d, N, T, tgt, hist, dict(snooker=.1, history_thin=10), seed=23)


def test_c5_shape_matches_oracle():
    """BASELINE config 5's shape at 8192 chains: 50-D correlated Gaussian, crossover adaptation during a 30-iteration
    burn-in (the adapted probabilities are compared as well), 60 iterations."""
    d, N, T = 50, 8192, 60
    rng = np.random.default_rng(145)
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(2 * N, d))
    _oracle_parity(d, N, T, tgt, hist, dict(snooker=.1, history_thin=10, adapt_crossover=True, crossover_burnin=30), seed=24)


def _check_invariants(eng, trace, logp, dec, tgt, nseed, thin, starts):
    d, N = eng.d, eng.N
    tr = trace[:, :, :d].cpu().numpy()                       # [N, T, d]
    lp = logp.cpu().numpy()
    dd = decode_decisions(dec.cpu().numpy().astype(np.uint32))
    T = tr.shape[1]
    # record_history (Dream.py:919-938): append #w holds every chain's state of iteration w*thin, in chain order
    Z = eng.Z[:eng.archive_rows, :d].cpu().numpy()
    app = [t for t in range(T) if t % thin == 0]
    assert eng.archive_rows == nseed + len(app) * N
    for w, t in enumerate(app):
        np.testing.assert_array_equal(Z[nseed + w * N: nseed + (w + 1) * N], tr[:, t, :])
    # a step that did not change the state repeats it (and its log-posterior) bit for bit
    prev = np.concatenate([starts[:, None, :], tr[:, :-1, :]], axis=1)
    same = (tr == prev).all(axis=2)
    np.testing.assert_array_equal(same, dd['changed'] == 0)
    assert np.array_equal(lp[:, 1:][same[:, 1:]], lp[:, :-1][same[:, 1:]])
    # log_ps = log-likelihood of the returned point (flat prior): spot-check chains with the host callable
    for c in (0, N // 2, N - 1):
        for t in (0, T // 2, T - 1):
            ref = tgt(tr[c, t])
            assert abs(lp[c, t] - ref) <= 10 * logp_tol(ref), (c, t, lp[c, t], ref)
    acc = dd['changed'].mean()
    assert 0.01 < acc < 0.9


def test_c2_full_size_window_kernel():
    from pydream_b200.engine import DreamEngine
    d, N, T, thin, nseed = 100, 1024, 120, 10, 4096
    rng = np.random.default_rng(42)
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(nseed, d))
    kw = dict(seed=17, snooker=.1, history_thin=thin)
    eng = DreamEngine(d, N, hist, hist[:N], tgt, **kw)
    trace, logp, dec = eng.run(T)
    _check_invariants(eng, trace, logp, dec, tgt, nseed, thin, hist[:N])
    # the generic kernel (direct quadratic form) takes the same decisions
    gen = DreamEngine(d, N, hist, hist[:N], tgt, generic_kernel=True, **kw)
    t2, l2, d2 = gen.run(T)
    assert bool((dec == d2).all().item())
    err = (logp - l2).abs().cpu().numpy()
    assert np.all(err <= logp_tol(l2.cpu().numpy()))


def test_c3_full_size_multitry_mixture():
    from pydream_b200.engine import DreamEngine
    d, N, T, thin, nseed = 10, 4096, 40, 10, 16384
    rng = np.random.default_rng(43)
    tgt = targets.BimodalMixture.benchmark(d)
    hist = rng.normal(size=(nseed, d))
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=5, snooker=.1, history_thin=thin, multitry=5)
    trace, logp, dec = eng.run(T)
    _check_invariants(eng, trace, logp, dec, tgt, nseed, thin, hist[:N])
    sel = decode_decisions(dec.cpu().numpy().astype(np.uint32))['sel']
    assert sel.min() >= 0 and sel.max() <= 4 and len(np.unique(sel)) == 5        # every multi-try slot gets picked


def test_c5_shape_adaptation_and_rhat():
    """d=50 Gaussian with crossover adaptation during burn-in, then Gelman-Rubin on the device trace."""
    from pydream_b200.engine import DreamEngine
    d, N, T, thin, nseed = 50, 8192, 60, 10, 20000
    rng = np.random.default_rng(44)
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(nseed, d))
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=6, snooker=.1, history_thin=thin, adapt_crossover=True, crossover_burnin=30)
    trace, logp, dec = eng.run(T)
    _check_invariants(eng, trace, logp, dec, tgt, nseed, thin, hist[:N])
    cr = eng.cr_probs.cpu().numpy()
    assert abs(cr.sum() - 1) < 1e-12 and np.all(cr > 0) and not np.allclose(cr, 1 / 3)
    rhat = eng.gelman_rubin(trace).cpu().numpy()
    tr = trace[:, T // 2:, :d].cpu().numpy()
    W = tr.var(axis=1).mean(axis=0)
    B = tr.mean(axis=1).var(axis=0)
    np.testing.assert_allclose(rhat, np.sqrt((W * (1 - 1. / T) + B) / W), rtol=1e-10)


def test_c4_shape_banana_200d():
    """200-D twisted Gaussian (two chunk rounds per lane), the per-GPU share of config 4 at 2 GPUs."""
    from pydream_b200.engine import DreamEngine
    d, N, T, thin, nseed = 200, 4096, 30, 10, 16384
    rng = np.random.default_rng(45)
    tgt = targets.Banana(d, 0.1)
    hist = rng.normal(size=(nseed, d)) * np.sqrt(np.concatenate([[100.0], np.ones(d - 1)]))
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=7, snooker=.1, history_thin=thin)
    trace, logp, dec = eng.run(T)
    _check_invariants(eng, trace, logp, dec, tgt, nseed, thin, hist[:N])
