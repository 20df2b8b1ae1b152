"""CPU tests: the C ABI library loads and exports every symbol include/dreamzs.h declares (no compute
calls), the launch planner, and the reference's argument checks / messages in run_dream
(pydream/tests/test_dream.py:34-50)."""
import os
import re

import numpy as np
import pytest
from scipy.stats import norm, uniform

from pydream_b200 import _cabi, targets
from pydream_b200.core import run_dream
from pydream_b200.Dream import Dream
from pydream_b200.model import Model
from pydream_b200.parameters import SampledParam, FlatParam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'dreamzs.h')).read()
    declared = set(re.findall(r'^(?:int|int64_t|void)\s+(dreamzs_\w+)\s*\(', hdr, flags=re.M))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dreamzs_abi_version() == _cabi.ABI_VERSION


def test_struct_sizes_match_header():
    import ctypes as C
    assert C.sizeof(_cabi.Config) == 14 * 4 + 4 * 8 + 8
    assert C.sizeof(_cabi.State) == 22 * 8
    assert C.sizeof(_cabi.Trace) == 5 * 8
    assert C.sizeof(_cabi.Peers) == 8 + 8 * 8 + 8 * 8 + 8 + 8 + 8
    assert C.sizeof(_cabi.Adapt) == 8 + 8 + 13 * 8


def test_plan_segments():
    from pydream_b200.engine import plan_segments, appends_in
    segs = plan_segments(0, 25, 10, -1)
    assert segs == [(0, 1), (1, 10), (11, 10), (21, 4)]
    for t, n in segs:   # only the last iteration of a launch may append
        assert all((t + i) % 10 != 0 for i in range(n - 1))
    assert sum(n for _, n in segs) == 25
    assert appends_in(0, 25, 10) == 3 and appends_in(1, 9, 10) == 0 and appends_in(1, 10, 10) == 1
    segs = plan_segments(0, 30, 10, 12)   # burn-in: one iteration per launch up to iteration 12
    assert segs[:13] == [(t, 1) for t in range(13)] and segs[13:] == [(13, 8), (21, 9)]
    assert plan_segments(7, 5, 1, -1) == [(t, 1) for t in range(7, 12)]


def onedmodel():
    return [SampledParam(norm, loc=-2, scale=3)], targets.SumShift(1, 3.0)


def multidmodel():
    mu = np.array([-6.6, 3, 1.0, -.12])
    sd = np.array([.13, 5, .9, 1.0])
    return [SampledParam(norm, loc=mu, scale=sd)], targets.SumShift(4, 3.0)


def test_fail_with_one_chain():
    """pydream/tests/test_dream.py:34-38"""
    param, like = onedmodel()
    with pytest.raises(Exception, match='Dream should be run with at least '):
        run_dream(param, like, nchains=1)


def test_total_var_dimension_init():
    """pydream/tests/test_dream.py:40-50"""
    param, like = onedmodel()
    step = Dream(model=Model(likelihood=like, sampled_parameters=param), variables=param)
    assert step.total_var_dimension == 1
    param, like = multidmodel()
    step = Dream(model=Model(likelihood=like, sampled_parameters=param), variables=param)
    assert step.total_var_dimension == 4


def test_gamma_array():
    """pydream/tests/test_dream.py:68-76"""
    param, like = onedmodel()
    step = Dream(model=Model(likelihood=like, sampled_parameters=param), DEpairs=5)
    for d_prime, gamma_value in zip(range(1, 6), [1.683, 1.19, 0.972, 0.841, 0.753]):
        assert round(abs(step.gamma_arr[0][d_prime - 1][0] - gamma_value), 3) == 0


def test_restart_and_seed_checks():
    param, like = multidmodel()
    with pytest.raises(Exception, match='Restart run specified but no start positions given.'):
        run_dream(param, like, nchains=3, restart=True)
    with pytest.raises(Exception, match='no model name to load history'):
        run_dream(param, like, nchains=3, restart=True, start=[np.zeros(4)] * 3)
    with pytest.raises(Exception, match='The size of the seeded starting history is insufficient'):
        run_dream(param, like, nchains=30, nseedchains=10)
    with pytest.raises(Exception, match='not implemented yet'):
        run_dream([FlatParam(test_value=np.zeros(4))], like, nchains=3, niterations=5)


def test_dream_defaults_and_multitry_rules():
    param, like = multidmodel()
    m = Model(likelihood=like, sampled_parameters=param)
    s = Dream(model=m)
    assert (s.nCR, s.ngamma, s.multitry, s.snooker, s.p_gamma_unity, s.lamb, s.zeta, s.history_thin) == \
        (3, 1, 1, .1, .2, .05, 1e-12, 10)
    assert s.nseedchains == 40 and list(s.CR_values) == [1 / 3., 2 / 3., 1.0]
    assert Dream(model=m, multitry=True).multitry == 5 and Dream(model=m, multitry=7).multitry == 7
    assert Dream(model=m, nCR=9).nCR == 4          # clipped to the dimension, as the reference does
    assert Dream(model=m, some_unknown_kwarg=1).extra_kwargs == {'some_unknown_kwarg': 1}


def test_closed_form_priors():
    p = SampledParam(uniform, loc=np.array([-5., 3.]), scale=np.array([15., 5.]))
    k, a, b = p.closed_form()
    assert k == 2 and list(a) == [-5., 3.] and list(b) == [15., 5.]
    assert SampledParam(norm, loc=-2, scale=3).closed_form()[0] == 1
    from scipy.stats import gamma
    assert SampledParam(gamma, 2.0).closed_form() is None
    x = np.array([0.5, 4.0])
    assert np.isclose(Model(targets.SumShift(2), [p]).total_logp(x)[0], -np.log(15.) - np.log(5.))


def test_targets_match_reference_example_formulas():
    d = 6
    g = targets.CorrelatedGaussian.benchmark(d)
    A = .5 * np.identity(d) + .5 * np.ones((d, d))
    Cm = np.array([[A[i][j] * np.sqrt((i + 1) * (j + 1)) for j in range(d)] for i in range(d)])
    x = np.linspace(-1, 2, d)
    ref = np.log(((2 * np.pi) ** (-d / 2)) * np.linalg.det(Cm) ** (-1. / 2)) - .5 * np.sum(x * np.dot(np.linalg.inv(Cm), x))
    assert np.isclose(g(x), ref, rtol=1e-13)
    mix = targets.BimodalMixture.benchmark(10)
    x = np.full(10, 5.0)
    assert np.isclose(mix(x), np.log(np.exp(-9.5949) + np.exp(-.5 * 1000 - 10.2880)))


def test_c_abi_argument_checks_without_a_gpu():
    """Entry points validate their arguments before touching the device: negative status codes, no exceptions."""
    import ctypes as C
    lib = _cabi.load()
    cfg = _cabi.Config(abi_version=_cabi.ABI_VERSION, ndim=4, ld=4, nchains_global=8, chain_begin=0, nchains_local=8, nCR=3,
                       ngamma=1, nDEpairs=1, multitry=1, hardboundaries=1, history_thin=10, target_kind=0, flags=0,
                       snooker=.1, p_gamma_unity=.2, lamb=.05, zeta=1e-12, seed=1)
    st, tr = _cabi.State(), _cabi.Trace()
    null_hook, null_adapt = _cabi.APPEND_HOOK(), None
    # null state pointers
    assert lib.dreamzs_step(C.byref(cfg), C.byref(st), C.byref(tr), 0, 1, 16, None) == _cabi.E_BADARG
    assert lib.dreamzs_init_logp(C.byref(cfg), C.byref(st), None) == _cabi.E_BADARG
    # wrong ABI version
    bad = _cabi.Config.from_buffer_copy(cfg)
    bad.abi_version = _cabi.ABI_VERSION + 1
    assert lib.dreamzs_step(C.byref(bad), C.byref(st), C.byref(tr), 0, 1, 16, None) == _cabi.E_BADARG
    # the native loop: negative counts, a sharded run without any transport for the other shards' rows
    assert lib.dreamzs_run(C.byref(cfg), C.byref(st), C.byref(tr), 0, -1, 16, 0, None, null_hook, None, null_adapt, None,
                           None, None) == _cabi.E_BADARG
    shard = _cabi.Config.from_buffer_copy(cfg)
    shard.nchains_local = 4
    assert lib.dreamzs_run(C.byref(shard), C.byref(st), C.byref(tr), 0, 5, 16, 0, None, null_hook, None, null_adapt, None,
                           None, None) == _cabi.E_BADARG
    # zero iterations is a no-op that reports the unchanged archive size
    rows, nl = C.c_int64(-1), C.c_int64(-1)
    assert lib.dreamzs_run(C.byref(cfg), C.byref(st), C.byref(tr), 3, 0, 16, 0, None, null_hook, None, null_adapt, None,
                           C.byref(nl), C.byref(rows)) == _cabi.OK
    assert rows.value == 16 and nl.value == 0
    # pitched copy, split step, shared allocations
    assert lib.dreamzs_copy_d2h_2d(None, 8, None, 8, 8, 1, None) == _cabi.E_BADARG
    assert lib.dreamzs_propose(C.byref(cfg), C.byref(st), 0, 16, None, None, None) == _cabi.E_BADARG
    assert lib.dreamzs_accept(C.byref(cfg), C.byref(st), C.byref(tr), 0, 16, None, None, None, None) == _cabi.E_BADARG
    assert lib.dreamzs_select(C.byref(cfg), C.byref(st), 0, 16, None, None, None, None, None) == _cabi.E_BADARG
    assert lib.dreamzs_shared_alloc(0, None, None) == _cabi.E_BADARG
    assert lib.dreamzs_gr_finish(None, None, 1, 1, 1, None, None) == _cabi.E_BADARG


def test_appends_in_matches_the_schedule():
    from pydream_b200.engine import plan_segments, appends_in
    for t0, n, thin in ((0, 25, 10), (1, 9, 10), (7, 40, 3), (5, 0, 4), (10, 1, 10), (11, 9, 10)):
        segs = plan_segments(t0, n, thin, -1)
        assert sum(m for _, m in segs) == n
        assert sum(1 for t, m in segs if (t + m - 1) % thin == 0) == appends_in(t0, n, thin)
        assert all((t + i) % thin != 0 for t, m in segs for i in range(m - 1))      # only a launch's last iteration appends


def test_run_dream_host_path_reaches_engine(monkeypatch):
    """Everything run_dream does on the host before the first launch (option validation, archive seed from the priors or
    from history_file, start positions, defaults such as crossover_burnin = niterations/10), observed at the engine's
    constructor (no GPU needed)."""
    import pydream_b200.engine as E

    class Reached(Exception):
        pass

    seen = {}

    class FakeEngine:
        def __init__(self, *a, **k):
            seen['a'], seen['k'] = a, k
            raise Reached()

    monkeypatch.setattr(E, 'DreamEngine', FakeEngine)
    param, like = multidmodel()
    for kw, want in ((dict(), dict(multitry=1, DEpairs=1, gamma_levels=1, adapt_gamma=False)),
                     (dict(tempering=True), dict(multitry=1)),
                     (dict(multitry=True, DEpairs=2, gamma_levels=3, adapt_gamma=True), dict(multitry=5, DEpairs=2, gamma_levels=3, adapt_gamma=True)),
                     (dict(hardboundaries=False, nCR=9), dict(hardboundaries=False, nCR=4))):
        with pytest.raises(Reached):
            run_dream(param, like, nchains=6, niterations=40, verbose=False, seed=1, **kw)
        a, k = seen['a'], seen['k']
        assert a[0] == 4 and a[1] == 6 and np.asarray(a[2]).shape == (40, 4) and np.asarray(a[3]).shape == (6, 4)
        assert k['crossover_burnin'] == 4 and k['history_thin'] == 10 and k['seed'] == 1
        for name, value in want.items():
            assert k[name] == value, (name, k[name])
        np.testing.assert_allclose(k['cr_probs'], np.full(k['nCR'], 1. / k['nCR']))
    d = 12
    hist = np.random.default_rng(0).uniform(-5, 15, size=(30, d))
    with pytest.raises(Reached):      # the analytic examples' way: FlatParam + history_file + explicit starts
        run_dream(FlatParam(test_value=np.zeros(d)), targets.CorrelatedGaussian.benchmark(d), nchains=5, niterations=20,
                  start=[hist[c] for c in range(5)], start_random=False, history_file=hist, verbose=False, seed=0)
    a, k = seen['a'], seen['k']
    np.testing.assert_array_equal(np.asarray(a[2]), hist)
    np.testing.assert_array_equal(np.asarray(a[3]), hist[:5])
    with pytest.raises(Reached):      # one start array for every chain (core.py:77-78)
        run_dream(FlatParam(test_value=np.zeros(d)), targets.CorrelatedGaussian.benchmark(d), nchains=5, niterations=20,
                  start=hist[3], start_random=False, history_file=hist, verbose=False)
    np.testing.assert_array_equal(np.asarray(seen['a'][3]), np.tile(hist[3], (5, 1)))


def test_dream_option_set_matches_reference(capsys):
    """Attributes and warning texts of pydream_b200.Dream against the reference's Dream.__init__ (pydream/Dream.py:63-191)
    for a few option sets; runs where the reference checkout exists."""
    import sys
    if not os.path.isdir('/root/reference/pydream'):
        pytest.skip('reference checkout not present')
    sys.path.insert(0, '/root/reference')
    import pydream.Dream as RD
    import pydream.model as RM
    import pydream.parameters as RP

    def build(SP, M, D, kw):
        params = [SP(norm, loc=np.arange(4.), scale=np.full(4, 1.5)), SP(uniform, loc=-1., scale=3.)]
        capsys.readouterr()
        obj = D(model=M(likelihood=lambda x: 0., sampled_parameters=params), **kw)
        return obj, capsys.readouterr().out

    for kw in (dict(), dict(nCR=9, gamma_levels=3, DEpairs=3, multitry=True, adapt_gamma=True, nseedchains=77, snooker=0, model_name='x'),
               dict(multitry=1), dict(multitry=0), dict(multitry=4, zeta=1e-3, history_thin=3, crossover_burnin=17)):
        r, rout = build(RP.SampledParam, RM.Model, RD.Dream, kw)
        o, oout = build(SampledParam, Model, Dream, kw)
        assert rout == oout
        for name in ('total_var_dimension', 'nCR', 'ngamma', 'njoint_cr_gamma_probs', 'crossover_burnin', 'adapt_crossover',
                     'adapt_gamma', 'snooker', 'p_gamma_unity', 'multitry', 'parallel', 'lamb', 'zeta', 'nseedchains',
                     'history_thin', 'start_random', 'save_history', 'history_file', 'verbose', 'model_name', 'iter',
                     'len_history', 'last_logp', 'gamma', 'chain_n', 'nchains', 'boundaries'):
            assert np.all(getattr(r, name) == getattr(o, name)), name
        for name in ('mins', 'maxs', 'CR_values', 'gamma_level_values', 'DEpairs', 'gamma_arr', 'CR_probabilities',
                     'gamma_probabilities', 'boundary_mask'):
            np.testing.assert_array_equal(np.asarray(getattr(r, name), dtype=float), np.asarray(getattr(o, name), dtype=float), err_msg=name)


def test_whitening_factor_and_packing():
    """Host side of the whitened window kernel: invC = L L^T to rounding, |L^T x|^2 agrees with x.(invC x) far inside the
    1e-12 tolerance, the packed tile order is the one include/dreamzs.h documents, a matrix that is not positive
    definite is refused."""
    from pydream_b200.engine import whitening_factor, pack_whitening
    rng = np.random.default_rng(0)
    for d in (7, 50, 100, 128):
        tgt = targets.CorrelatedGaussian.benchmark(d)
        Lw = whitening_factor(tgt.invC)
        assert Lw is not None and np.allclose(np.triu(Lw, 1), 0)
        assert np.abs(Lw @ Lw.T - tgt.invC).max() <= 1e-13 * np.abs(tgt.invC).max()
        for _ in range(20):
            x = rng.uniform(-5, 15, size=d)
            q_ref = np.sum(x * np.dot(tgt.invC, x))
            u = Lw.T @ x
            assert abs(np.dot(u, u) - q_ref) <= 1e-13 * abs(q_ref)
        ld = (d + 3) // 4 * 4
        packed = pack_whitening(Lw, ld)
        assert packed.size == int(_cabi.load().dreamzs_whiten_doubles(ld))
        nK = ld // 4
        tile0 = lambda I: I * nK - I * (I - 1)
        for (j, i) in ((0, 0), (d - 1, 0), (d - 1, d - 1), (d // 2, d // 3), (5, 6)):
            I, k = i // 8, j // 4
            got = packed[(tile0(I) + k - 2 * I) * 32 + (j % 4) + 4 * (i % 8)] if k >= 2 * I else 0.0
            assert got == (Lw[j, i] if j >= i else 0.0)
    bad = np.eye(4)
    bad[3, 3] = -1.0
    assert whitening_factor(bad) is None


def test_draw_ws_bytes_host_function():
    """dreamzs_draw_ws_bytes (include/dreamzs.h) is host arithmetic: records of the two-stage steps."""
    import ctypes as C
    lib = _cabi.load()

    def cfg(**kw):
        base = dict(abi_version=_cabi.ABI_VERSION, ndim=10, ld=12, nchains_global=4096, chain_begin=0, nchains_local=4096, nCR=3,
                    ngamma=1, nDEpairs=1, multitry=5, hardboundaries=1, history_thin=10, target_kind=2, flags=0, snooker=.1,
                    p_gamma_unity=.2, lamb=.05, zeta=1e-12, seed=1)
        base.update(kw)
        return _cabi.Config(**base)

    f = lambda c, n: int(lib.dreamzs_draw_ws_bytes(C.byref(c), n))
    # multi-try: 8 scalars + (2k-1) points x (A[ld] + B[ld]) doubles per (chain, iteration); C3: 1792 B
    assert f(cfg(), 10) == 4096 * 10 * (8 + 9 * 2 * 12) * 8
    assert f(cfg(multitry=8), 3) == 4096 * 3 * (8 + 15 * 2 * 12) * 8
    assert f(cfg(ndim=40, ld=40, multitry=4), 1) == 4096 * (8 + 7 * 2 * 40) * 8     # 8 lanes x 2 chunks per point, 4 points
    assert f(cfg(ndim=80, ld=80, multitry=3), 1) == 0            # a point needs more than 8 lanes x 2 chunks
    assert f(cfg(ndim=40, ld=40, multitry=5), 1) == 0            # five 8-lane points do not fit a warp
    assert f(cfg(flags=_cabi.FLAG_GENERIC_KERNEL), 10) == 0
    # single try: 4 + 2 ld doubles per (chain, iteration), whole windows while they fit the budget
    per = 8192 * (4 + 2 * 200) * 8
    assert f(cfg(ndim=200, ld=200, nchains_global=8192, nchains_local=8192, multitry=1, target_kind=3), 10) == 10 * per
    assert f(cfg(ndim=200, ld=200, nchains_global=8192, nchains_local=8192, multitry=1, target_kind=3), 1000) == (1024 * 2 ** 20 // per) * per
    assert f(cfg(multitry=1, target_kind=5), 10) == 0            # caller-evaluated likelihoods use the split step
    assert f(cfg(), 0) == 0
