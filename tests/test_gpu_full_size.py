"""Size-independent properties at BASELINE.json's full single-GPU sizes (the oracle is too slow there):
the archive tail is exactly the states at the appending iterations, rejected steps repeat the previous state,
the log-posterior trace is the target evaluated on the trace, and two independent kernels agree."""
import numpy as np
import pytest

from golden_util import logp_tol, decode_decisions
from pydream_b200 import targets

pytestmark = pytest.mark.gpu


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
