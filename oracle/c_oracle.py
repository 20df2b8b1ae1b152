"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): ctypes front-end of oracle/dreamzs_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_NCR, MAX_NGAMMA = 16, 8
ABI_VERSION = 4     # DREAMZS_ABI_VERSION of include/dreamzs.h (checked by dreamzs_oracle_run)


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'abi_version', 'ndim', 'ld', 'nchains_global', 'chain_begin', 'nchains_local', 'nCR', 'ngamma', 'nDEpairs',
        'multitry', 'hardboundaries', 'history_thin', 'target_kind', 'flags')] + [
        ('snooker', C.c_double), ('p_gamma_unity', C.c_double), ('lamb', C.c_double), ('zeta', C.c_double),
        ('seed', C.c_uint64)]


class State(C.Structure):
    _fields_ = [('Z', C.c_void_p), ('Z_capacity_rows', C.c_int64), ('X', C.c_void_p), ('last_prior', C.c_void_p),
                ('last_like', C.c_void_p), ('cr_probs', C.c_void_p), ('gamma_probs', C.c_void_p),
                ('gamma_table', C.c_void_p), ('target_table', C.c_void_p), ('prior_kind', C.c_void_p),
                ('prior_a', C.c_void_p), ('prior_b', C.c_void_p), ('mins', C.c_void_p), ('maxs', C.c_void_p),
                ('gauss_Y', C.c_void_p), ('gauss_Q', C.c_void_p), ('gauss_L', C.c_void_p), ('gauss_U', C.c_void_p),
                ('sync_ws', C.c_void_p), ('sync_ws_words', C.c_int64),
                ('draw_ws', C.c_void_p), ('draw_ws_bytes', C.c_int64)]     # the last four: unused by the oracle (NULL)


class Adapt(C.Structure):
    _fields_ = [('adapt_crossover', C.c_int32), ('adapt_gamma', C.c_int32), ('crossover_burnin', C.c_int64),
                ('cr_probs', C.c_void_p), ('ncr_updates', C.c_void_p), ('delta_m', C.c_void_p),
                ('gamma_probs', C.c_void_p), ('ngamma_updates', C.c_void_p), ('delta_m_gamma', C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, 'libdreamzs_oracle.so')
    srcs = [os.path.join(_HERE, 'dreamzs_oracle.c'), os.path.join(_HERE, '..', 'include', 'dreamzs.h')]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(['make', '-s', '-C', _HERE, '-B', 'libdreamzs_oracle.so'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.dreamzs_oracle_run.restype = C.c_int
        _LIB.dreamzs_oracle_run_pt.restype = C.c_int
        _LIB.dreamzs_oracle_work_doubles.restype = C.c_int64
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def round_up4(d):
    return (int(d) + 3) // 4 * 4


def gamma_table(ngamma, nDEpairs, ndim):
    out = np.zeros((ngamma, nDEpairs, ndim))
    lib().dreamzs_oracle_gamma_table(C.c_int32(ngamma), C.c_int32(nDEpairs), C.c_int32(ndim), _p(out))
    return out


def gelman_rubin(trace):
    """trace: (nchains, nsamples, d) contiguous."""
    trace = np.ascontiguousarray(trace, dtype=np.float64)
    n, s, d = trace.shape
    out = np.zeros(d)
    lib().dreamzs_oracle_gelman_rubin(_p(trace), C.c_int64(n), C.c_int64(s), C.c_int32(d), C.c_int64(d), _p(out))
    return out


def temperature_ladder(nchains):
    """T[i] = 0.001 ** (i / nchains), pydream/core.py:133-136."""
    return np.array([np.power(.001, (float(i) / nchains)) for i in range(nchains)])


class OracleSampler:
    """Lock-step CPU sampler with the option names of Dream.__init__ (pydream/Dream.py:63-67)."""

    def __init__(self, ndim, nchains, history, starts, target_kind, target_table, seed=0, nCR=3, gamma_levels=1,
                 DEpairs=1, multitry=1, snooker=.1, p_gamma_unity=.2, lamb=.05, zeta=1e-12, history_thin=10,
                 hardboundaries=True, adapt_crossover=False, adapt_gamma=False, crossover_burnin=0,
                 prior_kind=None, prior_a=None, prior_b=None, capacity_rows=None, cr_probs=None, gamma_probs=None,
                 nthreads=1, chain_begin=0, nchains_global=None):
        """nchains / starts describe the chains stepped HERE; a shard passes chain_begin and nchains_global (SURVEY.md
        8(e): global chain ids in the random streams, archive row of an append = rows so far + global chain id) and
        fetches the other shards' appended rows itself after every appending iteration (Z is exposed)."""
        d, N = int(ndim), int(nchains)
        self.Ng = int(nchains_global) if nchains_global is not None else N
        self.d, self.N, self.ld = d, N, round_up4(d)
        history = np.asarray(history, dtype=np.float64).reshape(-1, d)
        self.nseed = history.shape[0]
        cap = int(capacity_rows) if capacity_rows else self.nseed
        self.Z = np.zeros((max(cap, self.nseed), self.ld))
        self.Z[:self.nseed, :d] = history
        self.X = np.zeros((N, self.ld))
        self.X[:, :d] = np.asarray(starts, dtype=np.float64).reshape(N, d)
        self.last_prior, self.last_like = np.zeros(N), np.zeros(N)
        self.count = C.c_int64(0)
        self.nthreads = int(nthreads)
        self.cfg = Config(abi_version=ABI_VERSION, ndim=d, ld=self.ld, nchains_global=self.Ng, chain_begin=int(chain_begin), nchains_local=N,
                          nCR=nCR, ngamma=gamma_levels, nDEpairs=DEpairs, multitry=multitry,
                          hardboundaries=int(bool(hardboundaries)), history_thin=history_thin,
                          target_kind=int(target_kind), snooker=snooker, p_gamma_unity=p_gamma_unity, lamb=lamb,
                          zeta=zeta, seed=seed)
        self.cr_probs = np.array(cr_probs if cr_probs is not None else [1 / float(nCR)] * nCR, dtype=np.float64)
        self.gamma_probs = np.array(gamma_probs if gamma_probs is not None else [1 / float(gamma_levels)] * gamma_levels,
                                    dtype=np.float64)
        self.ncr_updates, self.delta_m = np.zeros(nCR), np.zeros(nCR)
        self.ngamma_updates, self.delta_m_gamma = np.zeros(gamma_levels), np.zeros(gamma_levels)
        self.gamma_table = gamma_table(gamma_levels, DEpairs, d)
        self.target_table = np.ascontiguousarray(target_table, dtype=np.float64)
        self.prior_kind = np.zeros(d, dtype=np.int32) if prior_kind is None else np.ascontiguousarray(prior_kind, dtype=np.int32)
        self.prior_a = np.zeros(d) if prior_a is None else np.ascontiguousarray(prior_a, dtype=np.float64)
        self.prior_b = np.ones(d) if prior_b is None else np.ascontiguousarray(prior_b, dtype=np.float64)
        self.mins = np.full(d, -np.inf)
        self.maxs = np.full(d, np.inf)
        u = self.prior_kind == 2
        self.mins[u] = self.prior_a[u]
        self.maxs[u] = self.prior_a[u] + self.prior_b[u]
        self.st = State(Z=_p(self.Z), Z_capacity_rows=self.Z.shape[0], X=_p(self.X), last_prior=_p(self.last_prior),
                        last_like=_p(self.last_like), cr_probs=_p(self.cr_probs), gamma_probs=_p(self.gamma_probs),
                        gamma_table=_p(self.gamma_table), target_table=_p(self.target_table),
                        prior_kind=_p(self.prior_kind), prior_a=_p(self.prior_a), prior_b=_p(self.prior_b),
                        mins=_p(self.mins), maxs=_p(self.maxs))
        self.ad = Adapt(adapt_crossover=int(bool(adapt_crossover)), adapt_gamma=int(bool(adapt_gamma)),
                        crossover_burnin=int(crossover_burnin), cr_probs=_p(self.cr_probs),
                        ncr_updates=_p(self.ncr_updates), delta_m=_p(self.delta_m), gamma_probs=_p(self.gamma_probs),
                        ngamma_updates=_p(self.ngamma_updates), delta_m_gamma=_p(self.delta_m_gamma))
        self.iter = 0
        lib().dreamzs_oracle_init_logp(C.byref(self.cfg), C.byref(self.st))

    def ensure_capacity(self, niter):
        thin = self.cfg.history_thin
        appends = sum(1 for t in range(self.iter, self.iter + niter) if t % thin == 0)
        need = self.nseed + self.count.value + appends * self.Ng
        if need > self.Z.shape[0]:
            Z = np.zeros((need, self.ld))
            Z[:self.Z.shape[0]] = self.Z
            self.Z = Z
            self.st.Z = _p(self.Z)
            self.st.Z_capacity_rows = need

    def run(self, niter, rows_dbg_n=0):
        """Returns dict(states (T,N,d), logp (T,N), decisions (T,N), rows (T,N,rows_dbg_n))."""
        niter = int(niter)
        self.ensure_capacity(niter)
        trace = np.zeros((self.N, niter, self.ld))
        logp = np.zeros((self.N, niter))
        dec = np.zeros((self.N, niter), dtype=np.uint32)
        rows = np.zeros((self.N, niter, rows_dbg_n), dtype=np.int64) if rows_dbg_n else None
        rc = lib().dreamzs_oracle_run(C.byref(self.cfg), C.byref(self.st), C.byref(self.ad), C.c_int64(self.iter),
                                      C.c_int64(niter), C.c_int64(self.nseed), C.byref(self.count), _p(trace), _p(logp),
                                      _p(dec), _p(rows), C.c_int32(rows_dbg_n), C.c_int32(self.nthreads))
        if rc != 0:
            raise RuntimeError('dreamzs_oracle_run failed: %d' % rc)
        self.iter += niter
        out = dict(states=np.ascontiguousarray(trace[:, :, :self.d].transpose(1, 0, 2)), logp=logp.T.copy(),
                   decisions=dec.T.copy())
        if rows is not None:
            out['rows'] = rows.transpose(1, 0, 2).copy()
        return out

    def run_pt(self, niter, temperature=None):
        """Parallel tempering (pydream/core.py:131-236).  Returns dict(sampled_params (N, 2 niter, d), log_ps (N, 2 niter),
        decisions (N, 2 niter), swaps (niter, 3) = first chain, second chain, accepted)."""
        niter = int(niter)
        self.ensure_capacity(niter)
        T = temperature_ladder(self.N) if temperature is None else np.ascontiguousarray(temperature, dtype=np.float64)
        trace = np.zeros((self.N, 2 * niter, self.ld))
        logp = np.zeros((self.N, 2 * niter))
        dec = np.zeros((self.N, 2 * niter), dtype=np.uint32)
        swaps = np.zeros((niter, 3), dtype=np.int64)
        rc = lib().dreamzs_oracle_run_pt(C.byref(self.cfg), C.byref(self.st), C.byref(self.ad), _p(T), C.c_int64(self.iter),
                                         C.c_int64(niter), C.c_int64(self.nseed), C.byref(self.count), _p(trace), _p(logp),
                                         _p(dec), _p(swaps), C.c_int32(self.nthreads))
        if rc != 0:
            raise RuntimeError('dreamzs_oracle_run_pt failed: %d' % rc)
        self.iter += niter
        return dict(sampled_params=np.ascontiguousarray(trace[:, :, :self.d]), log_ps=logp, decisions=dec, swaps=swaps)

    @property
    def history_flat(self):
        n = self.nseed + self.count.value
        return self.Z[:n, :self.d].reshape(-1).copy()
