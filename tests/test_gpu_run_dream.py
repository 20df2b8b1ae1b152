"""run_dream drop-in behaviour on the GPU, mirroring the reference's own integration tests
(pydream/tests/test_dream.py:397-497, 629-705): return shapes, archive file layout and length,
archive tail == multiset of returned samples at history_thin=1, restart files, hard boundaries."""
import os

import numpy as np
import pytest
from scipy.stats import norm, uniform

from pydream_b200 import targets
from pydream_b200.core import run_dream
from pydream_b200.convergence import Gelman_Rubin
from pydream_b200.parameters import SampledParam, FlatParam

pytestmark = pytest.mark.gpu


def multidmodel():
    mu = np.array([-6.6, 3, 1.0, -.12])
    sd = np.array([.13, 5, .9, 1.0])
    return [SampledParam(norm, loc=mu, scale=sd)], targets.SumShift(4, 3.0)


def multidmodel_uniform():
    lower = np.array([-5, -9, 5, 3])
    upper = np.array([10, 2, 7, 8])
    return [SampledParam(uniform, loc=lower, scale=upper - lower)], targets.SumShift(4, 3.0)


def test_return_shapes_and_history_file(tmp_path, monkeypatch):
    """test_history_correct_after_sampling_simple_model (test_dream.py:629-646) + length formula (:648-668)."""
    monkeypatch.chdir(tmp_path)
    params, like = multidmodel()
    nchains, niter = 5, 20
    sampled, logps = run_dream(params, like, niterations=niter, nchains=nchains, multitry=False, parallel=False,
                               history_thin=1, model_name='test_history_correct', adapt_crossover=False, verbose=False, seed=3)
    assert len(sampled) == nchains and sampled[0].shape == (niter, 4) and logps[0].shape == (niter, 1)
    history = np.load('test_history_correct_DREAM_chain_history.npy')
    nseed = 40
    assert history.ndim == 1 and len(history) == 4 * (nchains * niter + nseed)     # flat float64, rows of ndim
    tail = history[nseed * 4:].reshape(-1, 4)
    samples = np.concatenate(sampled)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    np.testing.assert_array_equal(key(tail), key(samples))
    assert os.path.exists('test_history_correct_DREAM_chain_adapted_crossoverprob.npy')
    assert os.path.exists('test_history_correct_DREAM_chain_adapted_gammalevelprob.npy')
    # log_ps are log_like + log_prior of the returned points (core.py:115)
    x = sampled[2][-1]
    ref = norm(loc=[-6.6, 3, 1.0, -.12], scale=[.13, 5, .9, 1.0]).logpdf(x).sum() + np.sum(x + 3)
    assert abs(logps[2][-1, 0] - ref) <= 1e-12 * max(1, abs(ref))


def test_history_length_with_thinning(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    params, like = multidmodel()
    run_dream(params, like, niterations=30, nchains=3, history_thin=10, model_name='thin', verbose=False, seed=1)
    history = np.load('thin_DREAM_chain_history.npy')
    assert len(history) == 4 * ((3 * 30) // 10 + 40)


def test_restart_uses_saved_files(tmp_path, monkeypatch):
    """restart=True reloads <model_name>_DREAM_chain_history.npy and the adapted probabilities
    (core.py:46-62, 255-263)."""
    monkeypatch.chdir(tmp_path)
    params, like = multidmodel()
    s1, _ = run_dream(params, like, niterations=40, nchains=3, history_thin=2, model_name='rs', verbose=False, seed=5)
    h1 = np.load('rs_DREAM_chain_history.npy')
    cr1 = np.load('rs_DREAM_chain_adapted_crossoverprob.npy')
    assert abs(cr1.sum() - 1) < 1e-12
    starts = [s1[c][-1] for c in range(3)]
    s2, _ = run_dream(params, like, niterations=20, nchains=3, history_thin=2, model_name='rs', verbose=False,
                      restart=True, start=starts, seed=6)
    h2 = np.load('rs_DREAM_chain_history.npy')
    assert len(h2) == len(h1) + 4 * (3 * 20 // 2)
    np.testing.assert_array_equal(h2[:len(h1)], h1)
    assert s2[0].shape == (20, 4)


def test_boundaries_obeyed_aftersampling():
    """test_dream.py:670-705: uniform priors + hardboundaries -> no sample outside the box."""
    params, like = multidmodel_uniform()
    sampled, logps = run_dream(params, like, niterations=1000, nchains=5, verbose=False, save_history=False, seed=2)
    lower, upper = np.array([-5, -9, 5, 3]), np.array([10, 2, 7, 8])
    for ch in sampled:
        assert np.all(ch >= lower) and np.all(ch <= upper)
    assert np.all(np.isfinite(np.concatenate(logps)))


def test_start_handling_and_flat_prior():
    """FlatParam needs history_file + start + start_random=False (both analytic examples do this,
    dream_ex_ndim_gaussian.py:65); a single start array is broadcast to every chain (core.py:77-78)."""
    d = 10
    rng = np.random.default_rng(0)
    hist = rng.normal(size=(100, d))
    like = targets.BimodalMixture.benchmark(d)
    params = FlatParam(test_value=np.zeros(d))
    sampled, logps = run_dream(params, like, niterations=50, nchains=3, start=[hist[c] for c in range(3)],
                               start_random=False, history_file=hist, multitry=5, verbose=False, save_history=False, seed=8)
    assert len(sampled) == 3 and sampled[0].shape == (50, d)
    s2, _ = run_dream(params, like, niterations=5, nchains=3, start=hist[0], start_random=False, history_file=hist,
                      verbose=False, save_history=False, seed=8)
    assert s2[1].shape == (5, d)
    rhat = Gelman_Rubin(sampled)
    assert rhat.shape == (d,) and np.all(np.isfinite(rhat))


def test_converges_to_gaussian_target():
    """Distributional sanity on the 10-D standard-normal prior with a flat likelihood (BASELINE config 1)."""
    d = 10
    params = [SampledParam(norm, loc=np.zeros(d), scale=np.ones(d))]
    sampled, _ = run_dream(params, targets.Constant(d, 0.0), niterations=3000, nchains=64, verbose=False,
                           save_history=False, seed=11, nseedchains=256)
    x = np.concatenate([s[1500:] for s in sampled])
    assert np.all(np.abs(x.mean(axis=0)) < 0.1) and np.all(np.abs(x.std(axis=0) - 1) < 0.1)
    assert np.all(Gelman_Rubin(sampled) < 1.1)


def test_streamed_chunks_equal_one_chunk():
    """run_dream streams the samples to the host in chunks while sampling continues; the chunk size must not
    change anything (odd chunk sizes, chunks that do not divide the run, a row stride with padding)."""
    d = 10
    rng = np.random.default_rng(5)
    hist = rng.normal(size=(64, d))
    like = targets.BimodalMixture.benchmark(d)
    params = FlatParam(test_value=np.zeros(d))
    kw = dict(niterations=57, nchains=7, start=[hist[c] for c in range(7)], start_random=False, history_file=hist,
              verbose=False, save_history=False, seed=21, history_thin=4)
    ref_s, ref_l = run_dream(params, like, stream_chunk=1000, **kw)
    for chunk in (1, 7, 56):
        s, l = run_dream(params, like, stream_chunk=chunk, **kw)
        for c in range(7):
            np.testing.assert_array_equal(s[c], ref_s[c])
            np.testing.assert_array_equal(l[c], ref_l[c])
    # adaptation during burn-in crosses chunk boundaries
    params2, like2 = multidmodel()
    hist2 = rng.normal(size=(40, 4)) * np.array([.13, 5, .9, 1.0]) + np.array([-6.6, 3, 1.0, -.12])
    kw2 = dict(niterations=45, nchains=5, verbose=False, save_history=False, seed=4, adapt_crossover=True, crossover_burnin=30,
               start=[hist2[c] for c in range(5)], start_random=False, history_file=hist2)
    a_s, a_l = run_dream(params2, like2, stream_chunk=1000, **kw2)
    b_s, b_l = run_dream(params2, like2, stream_chunk=8, **kw2)
    for c in range(5):
        np.testing.assert_array_equal(a_s[c], b_s[c])
        np.testing.assert_array_equal(a_l[c], b_l[c])


def test_verbose_prints_acceptance(capsys):
    params, like = multidmodel()
    run_dream(params, like, niterations=30, nchains=3, verbose=True, nverbose=10, save_history=False, seed=1)
    out = capsys.readouterr().out
    assert 'acceptance rate' in out and 'Iteration:  20' in out


def simple_likelihood(param):
    """pydream/tests/test_models.py:46-50, as a plain Python callable (the reference's likelihood contract)."""
    return np.sum(param + 3)


@pytest.mark.parametrize('multitry', [False, 3])
def test_plain_python_likelihood(multitry):
    """run_dream with an ordinary host callable likelihood(param_vec) -> float (pydream/model.py:30), the reference's
    main use: the step runs on the GPU (propose / select / accept), the callable on the host.  Same draws as the
    in-kernel restatement of the same function, and the same run on the oracle."""
    from oracle import c_oracle
    params, like = multidmodel()
    rng = np.random.default_rng(5)
    nchains, niter = 5, 60
    hist = rng.normal(size=(50, 4)) * np.array([.13, 5, .9, 1.0]) + np.array([-6.6, 3, 1.0, -.12])
    kw = dict(niterations=niter, nchains=nchains, start=[hist[c] for c in range(nchains)], start_random=False,
              history_file=hist, verbose=False, save_history=False, seed=21, multitry=multitry)
    s_py, l_py = run_dream(params, simple_likelihood, **kw)
    s_an, l_an = run_dream(params, like, **kw)
    pk = np.ones(4, dtype=np.int32)
    orc = c_oracle.OracleSampler(4, nchains, hist, hist[:nchains].copy(), like.kind, like.table(), seed=21, prior_kind=pk,
                                 prior_a=np.array([-6.6, 3, 1.0, -.12]), prior_b=np.array([.13, 5, .9, 1.0]),
                                 adapt_crossover=True, crossover_burnin=niter // 10, multitry=multitry or 1).run(niter)
    for c in range(nchains):
        np.testing.assert_allclose(s_py[c], s_an[c], rtol=1e-12)
        np.testing.assert_allclose(l_py[c], l_an[c], rtol=1e-12)
        np.testing.assert_allclose(s_py[c], orc['states'][:, c, :], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(l_py[c][:, 0], orc['logp'][:, c], rtol=1e-11)


def test_seed_makes_prior_draws_reproducible():
    """No history file, random starts: the archive seed and the starts are prior draws on the host (Dream.py:203-225).
    With `seed=` they come from a generator keyed by it, so two runs agree bit for bit and a different seed differs."""
    params, like = multidmodel()
    kw = dict(niterations=40, nchains=5, verbose=False, save_history=False)
    a, la = run_dream(params, like, seed=77, **kw)
    b, lb = run_dream(params, like, seed=77, **kw)
    c, _ = run_dream(params, like, seed=78, **kw)
    for x, y in zip(a + la, b + lb):
        np.testing.assert_array_equal(x, y)
    assert not np.array_equal(a[0], c[0])
