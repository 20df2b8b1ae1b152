"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Deterministic lock-step execution of the UNMODIFIED reference (``/root/reference/pydream``)
with counter-based RNG injected.  Used only in the build container (the reference does not
travel to the GPU box) to (a) pin the C restatement ``oracle/dreamzs_oracle.c`` and (b) write
the golden vectors under ``tests/golden/`` (``tests/golden/make_golden.py``).

Mechanics (nothing in /root/reference is edited):
  * ``pydream.Dream.np`` is replaced by a proxy module whose ``.random`` is `_NpRandomShim`,
    ``pydream.Dream.random`` by `_PyRandomShim`, ``pydream.Dream.time`` by a no-op sleeper.
    Every RNG entry point Dream.py uses is served from ``oracle.philox.Stream`` keyed by
    (seed, chain, iteration) with running call numbers per primitive.
  * shared state is built in-process with the reference tests' own trick
    (pydream/tests/test_dream.py:507-508): ``pool = _setup_mp_dream_pool(...);
    pool._initializer(*pool._initargs)``.
  * one shallow copy of the Dream instance per chain (the reference pickles one into each
    worker, pydream/core.py:75-80); every iteration calls ``astep`` for chain 0..N-1.
  * ``record_history`` / ``estimate_crossover_probabilities`` / ``estimate_gamma_level_probs``
    of each copy are queued during the sweep and replayed after it in chain order, then every
    copy adopts the shared probabilities: synchronous DREAM(ZS) semantics (all chains of
    iteration t read the archive as it stood after iteration t-1).
  * the burn-in barrier (Dream.py:385-407) is satisfied by presetting the shared counter so
    each chain sees "all finished" on arrival.
  * parallel tempering (`run_lockstep_pt`): the reference's own driver ``_sample_dream_pt``
    (pydream/core.py:131-236) runs unmodified on a stand-in pool whose ``map`` steps the chains in
    order in this process; the driver's own draws (``np.random.choice(nchains, 2, replace=False)``
    and ``np.random.uniform()``, core.py:183, 195) are served from the stream of the pseudo-chain
    ``philox.DRIVER_CHAIN``.
"""
import copy
import os
import sys
import tempfile
import types

import numpy as np

from . import philox as px

def _find_reference():
    """The unmodified reference: the checkout in the build container, else the copy installed (unchanged, by pip
    --target) under baseline/_ref, which travels to the GPU box."""
    inst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')
    for root in ('/root/reference', inst):
        if os.path.isdir(os.path.join(root, 'pydream')):
            return root
    return '/root/reference'


REF_ROOT = _find_reference()


class _Ctx:
    stream = None          # philox.Stream of the chain-step being executed
    driver = None          # philox.Stream of the tempering driver's draws of the current iteration
    log = None             # dict collecting decisions of that chain-step


class _NpRandomShim:
    """Stands in for ``np.random`` inside pydream.Dream."""

    @staticmethod
    def multinomial(n, pvals):
        assert n == 1
        idx = _Ctx.stream.multinomial_index(pvals)
        _Ctx.log['multinomial'].append(idx)
        out = np.zeros(len(pvals), dtype=np.int64)
        out[idx] = 1
        return out

    @staticmethod
    def normal(loc, scale, size):
        return loc + scale * _Ctx.stream.normal_vec(int(size))

    @staticmethod
    def uniform(low=0.0, high=1.0, size=None):
        if size is None:
            u = _Ctx.stream.uniform53(px.ST_UNIFORM_SCAL)
            return low + (high - low) * u
        if isinstance(size, tuple):
            # a (k, d) request is served as k consecutive vector calls of length d
            k, d = size
            rows = [low + (high - low) * _Ctx.stream.uniform32_vec(px.ST_UNIFORM_VEC, d) for _ in range(k)]
            return np.array(rows)
        return low + (high - low) * _Ctx.stream.uniform32_vec(px.ST_UNIFORM_VEC, int(size))

    @staticmethod
    def rand(m):
        return _Ctx.stream.uniform32_vec(px.ST_RAND, int(m))

    @staticmethod
    def randint(low, high, size=1):
        assert low == 1 and size == 1
        return np.array([1 + _Ctx.stream.randint(high - low)])


class _PyRandomShim:
    """Stands in for the ``random`` module inside pydream.Dream."""

    @staticmethod
    def sample(population, k):
        rows = _Ctx.stream.sample(len(population), k)
        _Ctx.log['rows'].extend(rows)
        return rows


def _masked(ufunc):
    def f(*args, where=True, **kw):
        if where is True:
            return ufunc(*args, **kw)
        out = np.zeros(np.broadcast(*args).shape)
        ufunc(*args, out=out, where=where, **kw)
        return out[()] if out.ndim == 0 else out
    return f


class _NoSleep:
    @staticmethod
    def sleep(_):
        return None


def _import_reference():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pydream.Dream as D
    import pydream.core as C
    from pydream import Dream_shared_vars as SV
    from pydream.model import Model
    from pydream import parameters as P
    return D, C, SV, Model, P


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'pydream'))


def run_lockstep(make_parameters, likelihood, nchains, niterations, starts, history, seed, max_rows=32,
                 max_multinomial=24, **dream_kwargs):
    """Run the unmodified reference in lock-step.

    make_parameters(P) -> list of reference SampledParam/FlatParam objects (P = pydream.parameters)
    history: (nseed, d) array used as ``history_file``; starts: (nchains, d).
    Returns dict of arrays: states (T,N,d), log_prior (T,N), log_like (T,N), accept (T,N) u8,
    rows (T,N,max_rows) i64 (-1 padded), multinomial (T,N,max_multinomial) i8 (-1 padded),
    history_final (flat), cr_probs (T,nCR), gamma_probs (T,ngamma).
    """
    D, C, SV, Model, P = _import_reference()
    proxy = types.ModuleType('numpy_proxy')
    proxy.__dict__.update(np.__dict__)
    proxy.random = _NpRandomShim
    # np.log / np.divide with a ``where=`` mask leave masked outputs UNINITIALISED in the reference
    # (Dream.py:824, 829, 831, 835); the oracle pins those outputs to 0 (DESIGN.md "pinned undefined values").
    proxy.log = _masked(np.log)
    proxy.divide = _masked(np.divide)
    saved = (D.np, D.random, D.time)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix='dreamzs_oracle_')
    try:
        os.chdir(tmp)
        D.np, D.random, D.time = proxy, _PyRandomShim, _NoSleep
        params = make_parameters(P)
        model = Model(likelihood=likelihood, sampled_parameters=params)
        hist_path = os.path.join(tmp, 'seed_history.npy')
        np.save(hist_path, np.asarray(history, dtype=np.float64))
        kw = dict(start_random=False, save_history=False, verbose=False, history_file=hist_path)
        kw.update(dream_kwargs)
        proto = D.Dream(model=model, variables=params, **kw)
        d = proto.total_var_dimension
        start_list = [np.array(s, dtype=np.float64) for s in starts]
        pool = C._setup_mp_dream_pool(nchains, niterations, proto, start_pt=start_list)
        pool._initializer(*pool._initargs)
        pool.close()
        pool.join()
        burnin = proto.crossover_burnin
        chains = [copy.copy(proto) for _ in range(nchains)]
        queue = []

        def defer(obj, name):
            real = getattr(obj, name)

            def wrapper(*a, **kw):
                cp = lambda x: np.array(x, dtype=np.float64, copy=True) if isinstance(x, np.ndarray) else x
                queue.append((real, tuple(cp(x) for x in a), {k: cp(v) for k, v in kw.items()}))
                if name == 'estimate_crossover_probabilities':
                    return obj.CR_probabilities
                if name == 'estimate_gamma_level_probs':
                    return obj.gamma_probabilities
                return None
            setattr(obj, name, wrapper)

        for ch in chains:
            for name in ('record_history', 'estimate_crossover_probabilities', 'estimate_gamma_level_probs'):
                defer(ch, name)

        T, N = niterations, nchains
        out = dict(states=np.zeros((T, N, d)), log_prior=np.zeros((T, N)), log_like=np.zeros((T, N)),
                   accept=np.zeros((T, N), dtype=np.uint8),
                   rows=-np.ones((T, N, max_rows), dtype=np.int64),
                   multinomial=-np.ones((T, N, max_multinomial), dtype=np.int8),
                   cr_probs=np.zeros((T, proto.nCR)), gamma_probs=np.zeros((T, proto.ngamma)))
        X = [s.copy() for s in start_list]
        for t in range(T):
            for c in range(N):
                _Ctx.stream = px.Stream(seed, c, t)
                _Ctx.log = dict(multinomial=[], rows=[])
                if t == burnin:
                    SV.nchains.value = N - 1
                q0 = X[c]
                q_new, lpri, llik = chains[c].astep(q0)
                q_new = np.array(q_new, dtype=np.float64).reshape(-1)
                out['accept'][t, c] = 0 if np.array_equal(q0, q_new) else 1
                out['states'][t, c] = q_new
                out['log_prior'][t, c] = lpri
                out['log_like'][t, c] = llik
                r = _Ctx.log['rows'][:max_rows]
                out['rows'][t, c, :len(r)] = r
                m = _Ctx.log['multinomial'][:max_multinomial]
                out['multinomial'][t, c, :len(m)] = m
                X[c] = q_new.copy()
            for real, a, kw in queue:
                res = real(*a, **kw)
                if real.__name__ == 'estimate_crossover_probabilities':
                    real.__self__.CR_probabilities = res
                elif real.__name__ == 'estimate_gamma_level_probs':
                    real.__self__.gamma_probabilities = res
            queue.clear()
            if proto.adapt_crossover and t <= burnin:
                shared = list(SV.cross_probs[0:proto.nCR])
                for ch in chains:
                    ch.CR_probabilities = shared
            if proto.adapt_gamma and t <= burnin:
                shared = list(SV.gamma_level_probs[0:proto.ngamma])
                for ch in chains:
                    ch.gamma_probabilities = shared
            out['cr_probs'][t] = np.array(chains[0].CR_probabilities, dtype=np.float64)
            out['gamma_probs'][t] = np.array(chains[0].gamma_probabilities, dtype=np.float64)
        out['history_final'] = np.frombuffer(SV.history.get_obj()).copy()
        out['count_final'] = np.array(SV.count.value)
        out['crossover_burnin'] = np.array(burnin)
        out['ncr_updates'] = np.array(SV.ncr_updates[:])
        out['delta_m'] = np.array(SV.delta_m[:])
        return out
    finally:
        D.np, D.random, D.time = saved
        os.chdir(cwd)


class _DriverRandomShim:
    """Stands in for ``np.random`` inside pydream.core (the parent process of the tempering driver)."""
    log = None

    @staticmethod
    def choice(n, size, replace=True):
        assert size == 2 and replace is False
        pair = _Ctx.driver.sample(int(n), 2)           # first pick, second pick (np.random.choice returns them in draw order)
        _DriverRandomShim.log['pairs'].append(pair)
        return np.array(pair)

    @staticmethod
    def uniform():
        return _Ctx.driver.uniform53(px.ST_UNIFORM_SCAL)


def run_lockstep_pt(make_parameters, likelihood, nchains, niterations, starts, history, seed, **dream_kwargs):
    """Parallel tempering: the unmodified ``pydream.core._sample_dream_pt`` on a lock-step stand-in pool.

    Returns dict: sampled_params (N, 2*niter, d), log_ps (N, 2*niter, 1) exactly as the reference driver
    returns them, pairs (niter, 2) the chains proposed for a swap, T (N,) the temperature ladder,
    log_prior / log_like (niter, N) as returned by astep (before the swap), history_final, cr_probs.
    """
    D, C, SV, Model, P = _import_reference()
    proxy = types.ModuleType('numpy_proxy')
    proxy.__dict__.update(np.__dict__)
    proxy.random = _NpRandomShim
    proxy.log = _masked(np.log)
    proxy.divide = _masked(np.divide)
    cproxy = types.ModuleType('numpy_proxy_core')
    cproxy.__dict__.update(np.__dict__)
    cproxy.random = _DriverRandomShim
    saved = (D.np, D.random, D.time, C.np)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix='dreamzs_oracle_')
    try:
        os.chdir(tmp)
        D.np, D.random, D.time, C.np = proxy, _PyRandomShim, _NoSleep, cproxy
        params = make_parameters(P)
        model = Model(likelihood=likelihood, sampled_parameters=params)
        hist_path = os.path.join(tmp, 'seed_history.npy')
        np.save(hist_path, np.asarray(history, dtype=np.float64))
        kw = dict(start_random=False, save_history=False, verbose=False, history_file=hist_path)
        kw.update(dream_kwargs)
        proto = D.Dream(model=model, variables=params, **kw)
        start_list = [np.array(s, dtype=np.float64) for s in starts]
        real_pool = C._setup_mp_dream_pool(nchains, niterations, proto, start_pt=start_list)
        real_pool._initializer(*real_pool._initargs)
        real_pool.close()
        real_pool.join()
        burnin = proto.crossover_burnin
        N = nchains
        extra = dict(log_prior=np.zeros((niterations, N)), log_like=np.zeros((niterations, N)),
                     cr_probs=np.zeros((niterations, proto.nCR)))
        _DriverRandomShim.log = dict(pairs=[])

        class LockstepPool:
            """pool.map(_sample_dream_pt_chain, args) of core.py:173: one astep per chain, in chain order; every worker
            of the reference holds its own (pickled) Dream instance, here one shallow copy per chain."""
            def __init__(self):
                self.chains = None
                self.queue = []
                self.t = 0

            def _defer(self, obj, name):
                real = getattr(obj, name)
                queue = self.queue

                def wrapper(*a, **kw):
                    cp = lambda x: np.array(x, dtype=np.float64, copy=True) if isinstance(x, np.ndarray) else x
                    queue.append((real, tuple(cp(x) for x in a), {k: cp(v) for k, v in kw.items()}))
                    if name == 'estimate_crossover_probabilities':
                        return obj.CR_probabilities
                    if name == 'estimate_gamma_level_probs':
                        return obj.gamma_probabilities
                    return None
                setattr(obj, name, wrapper)

            def map(self, fn, args):
                t = self.t
                if self.chains is None:
                    self.chains = [copy.copy(a[0]) for a in args]
                    for ch in self.chains:
                        for name in ('record_history', 'estimate_crossover_probabilities', 'estimate_gamma_level_probs'):
                            self._defer(ch, name)
                out = []
                for c, a in enumerate(args):
                    _Ctx.stream = px.Stream(seed, c, t)
                    _Ctx.log = dict(multinomial=[], rows=[])
                    if t == burnin:
                        SV.nchains.value = N - 1
                    res = fn((self.chains[c],) + tuple(a[1:]))
                    extra['log_prior'][t, c], extra['log_like'][t, c] = res[1], res[2]
                    out.append(res)
                for real, a, kw in self.queue:
                    res = real(*a, **kw)
                    if real.__name__ == 'estimate_crossover_probabilities':
                        real.__self__.CR_probabilities = res
                    elif real.__name__ == 'estimate_gamma_level_probs':
                        real.__self__.gamma_probabilities = res
                self.queue.clear()
                if proto.adapt_crossover and t <= burnin:
                    shared = list(SV.cross_probs[0:proto.nCR])
                    for ch in self.chains:
                        ch.CR_probabilities = shared
                if proto.adapt_gamma and t <= burnin:
                    shared = list(SV.gamma_level_probs[0:proto.ngamma])
                    for ch in self.chains:
                        ch.gamma_probabilities = shared
                extra['cr_probs'][t] = np.array(self.chains[0].CR_probabilities, dtype=np.float64)
                _Ctx.driver = px.Stream(seed, px.DRIVER_CHAIN, t)     # the swap draws that follow this map call
                self.t = t + 1
                return out

        sampled, log_ps = C._sample_dream_pt(nchains, niterations, proto, start_list, LockstepPool(), verbose=False)
        T = np.array([np.power(.001, (float(i) / nchains)) for i in range(nchains)])      # core.py:133-136
        out = dict(sampled_params=np.array(sampled), log_ps=np.array(log_ps),
                   pairs=np.array(_DriverRandomShim.log['pairs'], dtype=np.int64), T=T,
                   history_final=np.frombuffer(SV.history.get_obj()).copy(), count_final=np.array(SV.count.value),
                   crossover_burnin=np.array(burnin))
        out.update(extra)
        return out
    finally:
        D.np, D.random, D.time, C.np = saved
        os.chdir(cwd)
