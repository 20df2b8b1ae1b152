"""Split step for caller-evaluated likelihoods (dreamzs_propose -> torch callable -> dreamzs_accept; SURVEY.md 8(f)
row 2): with a torch restatement of an analytic target it must reproduce the C oracle (the checker pinned to the
unmodified reference) and the fused in-kernel step -- identical decisions and draws, log-posteriors within
1e-12 * max(1, |logp|)."""
import numpy as np
import pytest
from scipy.stats import uniform

from golden_util import logp_tol
from pydream_b200 import targets

pytestmark = pytest.mark.gpu


def _run(eng, T):
    trace, logp, dec = eng.run(T)
    d = eng.d
    return (trace[:, :, :d].permute(1, 0, 2).contiguous().cpu().numpy(), logp.t().contiguous().cpu().numpy(),
            dec.t().contiguous().cpu().numpy().astype(np.uint32))


def _oracle(d, N, hist, tgt, T, kw):
    """The same run on the C oracle -> (states, logp, decisions), iteration-major like _run."""
    from oracle import c_oracle
    kw = dict(kw)
    seed = kw.pop('seed')
    ref = c_oracle.OracleSampler(d, N, hist, hist[:N].copy(), tgt.kind, tgt.table(), seed=seed, nthreads=4, **kw).run(T)
    return ref['states'], ref['logp'], ref['decisions']


def _assert_matches_oracle(got, ref):
    np.testing.assert_array_equal(got[2], ref[2])
    assert np.all(np.abs(got[1] - ref[1]) <= logp_tol(ref[1])), (np.abs(got[1] - ref[1]) / logp_tol(ref[1])).max()
    np.testing.assert_allclose(got[0], ref[0], rtol=1e-10, atol=1e-11)


def _oracle_run_dream(lower, upper, hist, nchains, niter, seed, **kw):
    """What run_dream does with its defaults (crossover adaptation during the first niter/10 iterations) on the oracle:
    SumShift target, uniform prior on [lower, upper]."""
    from oracle import c_oracle
    tgt = targets.SumShift(len(lower), 3.0)
    pk = np.full(len(lower), 2, dtype=np.int32)
    orc = c_oracle.OracleSampler(len(lower), nchains, hist, hist[:nchains].copy(), tgt.kind, tgt.table(), seed=seed,
                                 prior_kind=pk, prior_a=lower, prior_b=upper - lower, adapt_crossover=True,
                                 crossover_burnin=niter // 10, **kw)
    return orc.run(niter)


def _torch_mixture(tgt):
    import torch
    mu0 = torch.as_tensor(tgt.table()[2:2 + tgt.ndim], device='cuda')
    mu1 = torch.as_tensor(tgt.table()[2 + tgt.ndim:2 + 2 * tgt.ndim], device='cuda')
    lf0, lf1 = float(tgt.table()[0]), float(tgt.table()[1])

    def fn(x):
        l0 = -.5 * ((x - mu0) ** 2).sum(dim=1) + lf0
        l1 = -.5 * ((x - mu1) ** 2).sum(dim=1) + lf1
        mx = torch.maximum(l0, l1)
        return torch.log(torch.exp(l0 - mx) + torch.exp(l1 - mx)) + mx
    return fn


@pytest.mark.parametrize('adapt', [False, True])
def test_external_mixture_equals_fused(adapt):
    from pydream_b200.engine import DreamEngine
    d, N, T = 10, 96, 48
    rng = np.random.default_rng(12)
    hist = rng.normal(size=(2 * N + 7, d)) * 3
    tgt = targets.BimodalMixture.benchmark(d)
    kw = dict(seed=31, snooker=.2, history_thin=4, adapt_crossover=adapt, crossover_burnin=30 if adapt else 0)
    a = _run(DreamEngine(d, N, hist, hist[:N], tgt, **kw), T)
    ext = targets.TorchLikelihood(d, _torch_mixture(tgt))
    eng = DreamEngine(d, N, hist, hist[:N], ext, **kw)
    b = _run(eng, 17)
    b2 = _run(eng, T - 17)          # a second call continues the same chains
    b = tuple(np.concatenate([b[i], b2[i]], axis=0) for i in range(3))
    _assert_matches_oracle(b, _oracle(d, N, hist, tgt, T, kw))      # the split step against the checker itself
    np.testing.assert_array_equal(a[2], b[2])
    assert np.all(np.abs(a[1] - b[1]) <= logp_tol(a[1])), (np.abs(a[1] - b[1]) / logp_tol(a[1])).max()
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12, atol=1e-13)


def test_external_with_bounded_prior_and_run_dream():
    """uniform prior + hard boundaries (reflection / redraw happen in dreamzs_propose), through run_dream."""
    import torch
    from pydream_b200.core import run_dream
    from pydream_b200.parameters import SampledParam
    lower, upper = np.array([-5., -9, 5, 3]), np.array([10., 2, 7, 8])
    params = [SampledParam(uniform, loc=lower, scale=upper - lower)]
    rng = np.random.default_rng(3)
    hist = rng.uniform(lower, upper, size=(64, 4))
    kw = dict(niterations=200, nchains=6, start=[hist[c] for c in range(6)], start_random=False, history_file=hist,
              verbose=False, save_history=False, seed=9)
    ref_s, ref_l = run_dream(params, targets.SumShift(4, 3.0), **kw)
    ext = targets.TorchLikelihood(4, lambda x: (x + 3.0).sum(dim=1))
    s, l = run_dream(params, ext, **kw)
    orc = _oracle_run_dream(lower, upper, hist, 6, 200, 9)
    for c in range(6):
        np.testing.assert_allclose(s[c], orc['states'][:, c, :], rtol=1e-10, atol=1e-11)
        assert np.all(np.abs(l[c][:, 0] - orc['logp'][:, c]) <= logp_tol(orc['logp'][:, c]))
        np.testing.assert_allclose(s[c], ref_s[c], rtol=1e-12)
        np.testing.assert_allclose(l[c], ref_l[c], rtol=1e-12)
        assert np.all(s[c] >= lower) and np.all(s[c] <= upper)
    # as a host callable (the reference's likelihood(param_vec) -> float contract)
    assert abs(ext(np.array([1., 2, 3, 4])) - 22.0) < 1e-12


@pytest.mark.parametrize('k,snooker', [(5, .2), (3, 0.)])
def test_external_multitry_equals_fused(k, snooker):
    """multi-try through dreamzs_propose -> fn -> dreamzs_select -> fn -> dreamzs_accept (Dream.py:275-323)."""
    from pydream_b200.engine import DreamEngine
    d, N, T = 10, 80, 36
    rng = np.random.default_rng(13)
    hist = rng.normal(size=(2 * N + 7, d)) * 3
    tgt = targets.BimodalMixture.benchmark(d)
    kw = dict(seed=32, snooker=snooker, history_thin=4, multitry=k)
    a = _run(DreamEngine(d, N, hist, hist[:N], tgt, **kw), T)
    eng = DreamEngine(d, N, hist, hist[:N], targets.TorchLikelihood(d, _torch_mixture(tgt)), **kw)
    b = _run(eng, T)
    eng.check_peers()
    _assert_matches_oracle(b, _oracle(d, N, hist, tgt, T, kw))      # the split step against the checker itself
    np.testing.assert_array_equal(a[2], b[2])
    assert np.all(np.abs(a[1] - b[1]) <= logp_tol(a[1])), (np.abs(a[1] - b[1]) / logp_tol(a[1])).max()
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12, atol=1e-13)


def test_external_multitry_bounded_prior():
    """multi-try + uniform prior: the boundary redraws (np.random.rand, Dream.py:749-775) keep their place in the stream
    across the three phases."""
    from pydream_b200.core import run_dream
    from pydream_b200.parameters import SampledParam
    lower, upper = np.array([-5., -9, 5, 3]), np.array([10., 2, 7, 8])
    params = [SampledParam(uniform, loc=lower, scale=upper - lower)]
    rng = np.random.default_rng(4)
    hist = rng.uniform(lower, upper, size=(64, 4))
    kw = dict(niterations=150, nchains=7, start=[hist[c] for c in range(7)], start_random=False, history_file=hist,
              verbose=False, save_history=False, seed=10, multitry=3)
    ref_s, ref_l = run_dream(params, targets.SumShift(4, 3.0), **kw)
    s, l = run_dream(params, targets.TorchLikelihood(4, lambda x: (x + 3.0).sum(dim=1)), **kw)
    orc = _oracle_run_dream(lower, upper, hist, 7, 150, 10, multitry=3)
    for c in range(7):
        np.testing.assert_allclose(s[c], orc['states'][:, c, :], rtol=1e-10, atol=1e-11)
        assert np.all(np.abs(l[c][:, 0] - orc['logp'][:, c]) <= logp_tol(orc['logp'][:, c]))
        np.testing.assert_allclose(s[c], ref_s[c], rtol=1e-12)
        np.testing.assert_allclose(l[c], ref_l[c], rtol=1e-12)


@pytest.mark.parametrize('name', ['const4_mt3_regen', 'sum10_mt5_regen'])
def test_external_multitry_regenerates_batches(name):
    """Uniform priors without hard boundaries: whole proposal batches fall outside the support (log prior -inf) and are
    regenerated (Dream.py:278-289) -- in the split step by dreamzs_repropose between the caller's evaluations."""
    import torch
    from oracle import c_oracle
    from pydream_b200.engine import DreamEngine
    from test_gpu_multitry import CASES, _case
    case = [c for c in CASES if c[0] == name][0]
    _, d, N, T, tgt, (pk, pa, pb), kw, hist = _case(case)
    kw = dict(kw, adapt_crossover=False)
    ref = c_oracle.OracleSampler(d, N, hist, hist[:N].copy(), tgt.kind, tgt.table(), seed=77, prior_kind=pk, prior_a=pa,
                                 prior_b=pb, **kw).run(T, rows_dbg_n=96)
    k, snk = kw['multitry'], (ref['decisions'] >> 1) & 1
    normal = np.where(snk == 1, 3 * (2 * k - 1), 2 * kw.get('DEpairs', 1) * (2 * k - 1))
    assert int(((ref['rows'] >= 0).sum(axis=2) > normal).sum()) > 0            # batches were in fact regenerated
    fn = (lambda x: torch.zeros(x.shape[0], dtype=torch.float64, device=x.device)) if name.startswith('const') \
        else (lambda x: (x + 3.0).sum(dim=1))
    eng = DreamEngine(d, N, hist, hist[:N].copy(), targets.TorchLikelihood(d, fn), pk, pa, pb, seed=77, **kw)
    got = _run(eng, T)
    eng.check_peers()
    np.testing.assert_array_equal(got[2], ref['decisions'])
    fin = np.isfinite(ref['logp'])
    assert np.array_equal(fin, np.isfinite(got[1]))
    assert np.all(np.abs(got[1][fin] - ref['logp'][fin]) <= 10 * logp_tol(ref['logp'][fin]))
    np.testing.assert_allclose(got[0], ref['states'], rtol=1e-10, atol=1e-11)
