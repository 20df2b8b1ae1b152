"""Device-resident MT-DREAM(ZS) engine: owns the HBM state and schedules the sm_100a kernels.

Replaces, for analytic targets, the reference's process pool + shared-memory namespace
(pydream/core.py:66-129, 250-327; pydream/Dream_shared_vars.py): every chain is a lane-group of
a step kernel, the archive ``Z`` lives in HBM, and the chain loop of ``_sample_dream`` becomes the native
loop ``dreamzs_run``: one persistent launch per span of windows for the dense Gaussian (whitened window kernel),
otherwise one launch -- or a draw kernel + a chain kernel -- per window of up to ``history_thin`` iterations (the
archive is immutable inside a window).  PyTorch is used only for device memory, streams and ``torch.distributed``.

Data layout in HBM (all float64, row stride ``ld`` = ndim rounded up to 4 doubles = 32 B):
    Z      [capacity_rows, ld]   archive, seed rows first, then N rows per append in chain order
    X      [N_local, ld]         current positions
    trace  [N_local, T, ld]      sampled_params (chain-major: chain c's trace is contiguous)
    logp   [N_local, T]          log_ps
    dec    [N_local, T] uint32   decision words (accept / snooker / CR / gamma level / multi-try pick)
Sharding: rank r owns global chains [r*N/G, (r+1)*N/G); Z is replicated on every GPU.  An appending iteration
stores each new row into ALL replicas (peer stores over NVLink from inside the step kernel; blocks confirmed by
system-scope flags, rows of not yet confirmed blocks read from the owner's archive: `dreamzs_peers`, DESIGN.md
section 6); only when the ranks cannot map each other's memory are the rows all-gathered in place with NCCL after
every appending launch instead.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _cabi
from . import targets as T
from .collectives import allgather_rows, allreduce_sum


GFLAG_STRIDE = 4096       # per-group flag words per peer rank (dreamzs_peers.gflag_stride)
# bytes in front of a shared archive block: flags (8 x uint64) at 0, error word at 1024, appended-chains counter at 2048,
# then from 8 * GFLAG_OFFSET on one row of GFLAG_STRIDE uint64 per peer rank (per-group append flags)
SHARED_HEADER = 8 * _cabi.GFLAG_OFFSET + _cabi.MAX_PEERS * GFLAG_STRIDE * 8


# Shared archive blocks (and the peers' mappings of them) kept between engines of the same process group: a run_dream call
# per model fit would otherwise pay cudaMalloc + handle exchange + cudaIpcOpenMemHandle for every peer + their inverses
# every time (tens to hundreds of ms at 4-8 GPUs).  An engine that needs no more than the cached block re-uses it after
# zeroing its header; release_shared_cache() frees them (collective).
_SHARED_CACHE = {}


def release_shared_cache(lib=None):
    """Collective over the groups that have cached blocks: unmap the peers' archives and free the shared blocks."""
    lib = lib or _cabi.load()
    for key in list(_SHARED_CACHE):
        ent = _SHARED_CACHE.pop(key)
        torch.cuda.synchronize(ent['device'])
        torch.distributed.barrier(ent['group'])
        for q, pq in enumerate(ent['opened']):
            if q != ent['rank']:
                lib.dreamzs_shared_close(C.c_void_p(pq))
        lib.dreamzs_shared_free(C.c_void_p(ent['base']))


class _DevicePtr:
    """Raw device memory (owned elsewhere) exposed through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr='|u1', data=(int(ptr), False), version=2)


def round_up4(d):
    return (int(d) + 3) // 4 * 4


def gamma_table(ngamma, nDEpairs, ndim):
    """Dream.gamma_arr (pydream/Dream.py:173-179), computed with numpy exactly as the reference does."""
    gamma_array = np.zeros((ngamma, nDEpairs, ndim))
    gamma_level_decrease = 1
    for gamma_level in range(1, ngamma + 1):
        for delta in range(1, nDEpairs + 1):
            gamma_array[gamma_level - 1, delta - 1, :] = (2.38 / np.sqrt(2 * delta * np.linspace(1, ndim, num=ndim))) / gamma_level_decrease
        gamma_level_decrease = gamma_level_decrease * 2
    return gamma_array


def device_target_table(target, ld):
    """Flat float64 table in the layout the kernels expect (include/dreamzs.h dreamzs_target_kind)."""
    if target.kind == T.TARGET_GAUSSIAN_DENSE:
        d = target.ndim
        At = np.zeros((d, ld))
        At[:, :d] = target.invC.T
        return np.concatenate([[target.log_F, 0.0], At.reshape(-1)])
    return np.ascontiguousarray(target.table(), dtype=np.float64)


def whitening_factor(invC):
    """Lower-triangular L with invC = L L^T (Cholesky of the symmetrised precision matrix in extended precision, rounded
    to float64), or None when the matrix is not safely positive definite.  The whitened window kernel carries u = L^T x
    and evaluates x.(invC x) as |u|^2 (include/dreamzs.h dreamzs_state.gauss_L)."""
    A = np.asarray(invC, dtype=np.longdouble)
    d = A.shape[0]
    S = (A + A.T) / 2
    Lf = np.zeros((d, d), dtype=np.longdouble)
    for j in range(d):
        s = S[j, j] - np.dot(Lf[j, :j], Lf[j, :j])
        if not np.isfinite(s) or s <= 0:
            return None
        Lf[j, j] = np.sqrt(s)
        if j + 1 < d:
            Lf[j + 1:, j] = (S[j + 1:, j] - Lf[j + 1:, :j] @ Lf[j, :j]) / Lf[j, j]
    L64 = Lf.astype(np.float64)
    back = L64.astype(np.longdouble) @ L64.astype(np.longdouble).T
    scale = float(np.abs(S).max())
    if not np.all(np.isfinite(L64)) or float(np.abs(back - S).max()) > 1e-13 * scale or float(np.abs(A - A.T).max()) > 1e-10 * scale:
        return None
    return L64


def pack_whitening(Lw, ld):
    """L in the tile order of dreamzs_state.gauss_L: for i-tile I and k = 2I .. ld/4-1 a tile of 32 doubles, entry
    (j % 4) + 4 (i % 8) = L[4k + j%4][8I + i%8]."""
    d = Lw.shape[0]
    nK, nI = ld // 4, (ld + 7) // 8
    P = np.zeros((8 * nI + 8, 8 * nI + 8))
    P[:d, :d] = Lw
    tiles = []
    for I in range(nI):
        for k in range(2 * I, nK):
            blk = P[4 * k:4 * k + 4, 8 * I:8 * I + 8]           # [j % 4, i % 8]
            tiles.append(blk.T.reshape(-1))                     # entry (i % 8) * 4 + (j % 4)
    return np.concatenate(tiles) if tiles else np.zeros(0)


def temperature_ladder(nchains):
    """T[i] = 0.001 ** (i / nchains): the ladder of _sample_dream_pt (pydream/core.py:133-136)."""
    T_ = np.zeros((nchains))
    for i in range(nchains):
        T_[i] = np.power(.001, (float(i) / nchains))
    return T_


def appends_in(iter_begin, niter, thin):
    """Number of iterations t in [iter_begin, iter_begin+niter) with t % thin == 0."""
    if niter <= 0:
        return 0
    first = ((iter_begin + thin - 1) // thin) * thin
    last = iter_begin + niter - 1
    return 0 if first > last else (last - first) // thin + 1


def plan_segments(iter_begin, niter, thin, single_until):
    """The launch schedule of the native loop (dreamzs_run in csrc/dreamzs_cabi.cu), restated for tests and tools:
    [iter_begin, iter_begin+niter) is split into launches; a launch ends at an appending iteration
    (t % thin == 0) and iterations t <= single_until run one per launch (burn-in adaptation)."""
    segs = []
    t, end = iter_begin, iter_begin + niter
    while t < end:
        if t <= single_until:
            n = 1
        else:
            nxt = ((t + thin - 1) // thin) * thin   # first appending iteration >= t
            n = min(end, nxt + 1) - t
        segs.append((t, n))
        t += n
    return segs


class DreamEngine:
    """Device state + launch schedule of one MT-DREAM(ZS) run (the GPU stand-in for the reference's pool of chain
    processes and its shared-memory namespace, pydream/core.py:250-327).

    history [nseed, ndim] seeds the archive, starts [nchains, ndim] are the chains' first states, `target` is one of
    pydream_b200.targets.  Keyword options carry the names and defaults of Dream.__init__ (pydream/Dream.py:63-67);
    `group` shards the chains over a torch.distributed process group (this rank owns a contiguous block),
    `reserve_iters` sizes the archive once for that many iterations, `peer_archive` selects NVLink peer stores
    (default) or the NCCL all-gather for keeping the archive replicas identical.  `run(n)` advances all chains by n
    iterations and returns device tensors; `run_to_host` streams the samples to pinned host memory meanwhile."""

    def __init__(self, ndim, nchains, history, starts, target, prior_kind=None, prior_a=None, prior_b=None, seed=0,
                 nCR=3, gamma_levels=1, DEpairs=1, multitry=1, snooker=.1, p_gamma_unity=.2, lamb=.05, zeta=1e-12,
                 history_thin=10, hardboundaries=True, adapt_crossover=False, adapt_gamma=False, crossover_burnin=0,
                 cr_probs=None, gamma_probs=None, device=None, group=None, record_decisions=True,
                 generic_kernel=False, window_kernel=True, reserve_iters=0, peer_archive=True, whitened=True, persistent=True, two_stage=True, draw_iters=None):
        if not torch.cuda.is_available():
            raise _cabi.DreamzsError('pydream_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = _cabi.load()
        # one launch for a whole span of windows (whitened window kernel); DREAMZS_NOPERSIST=1: one launch per window (A/B)
        self.persistent = bool(persistent) and not os.environ.get('DREAMZS_NOPERSIST')
        self.group = group
        self.world = torch.distributed.get_world_size(group) if group is not None else 1
        self.rank = torch.distributed.get_rank(group) if group is not None else 0
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        d, N = int(ndim), int(nchains)
        if N % self.world:
            raise ValueError('nchains (%d) must be divisible by the number of GPUs (%d)' % (N, self.world))
        self.d, self.N, self.ld = d, N, round_up4(d)
        self.Nl = N // self.world
        self.c0 = self.rank * self.Nl
        self.thin = int(history_thin)
        self.adapt_crossover, self.adapt_gamma = bool(adapt_crossover), bool(adapt_gamma)
        self.crossover_burnin = int(crossover_burnin)
        self.record_decisions = bool(record_decisions) or self.adapt_crossover or self.adapt_gamma
        self.target = target
        f64 = dict(dtype=torch.float64, device=self.device)
        history = np.asarray(history, dtype=np.float64).reshape(-1, d)
        self.nseed = history.shape[0]
        self.count = 0
        self.iter = 0
        self._hist_host = history
        self.Z = None
        # sharded runs keep the archive replicas identical with NVLink peer stores (dreamzs_peers) when the ranks can
        # map each other's memory, else with an NCCL all-gather per append
        self.peers = None
        self._shared = None          # (own base pointer, [opened peer base pointers])
        self.external = target.kind == T.TARGET_EXTERNAL     # likelihood evaluated by the caller: split step
        self._want_peers = bool(peer_archive) and self.world > 1 and self.world <= _cabi.MAX_PEERS and not self.external
        # the archive is sized once for `reserve_iters` iterations (it grows on demand beyond that)
        self._ensure_capacity(self.nseed + appends_in(0, int(reserve_iters), self.thin) * N)
        starts = np.asarray(starts, dtype=np.float64).reshape(N, d)
        Xh = np.zeros((self.Nl, self.ld))
        Xh[:, :d] = starts[self.c0:self.c0 + self.Nl]
        self.X = torch.from_numpy(Xh).to(self.device)
        self.last_prior = torch.zeros(self.Nl, **f64)
        self.last_like = torch.zeros(self.Nl, **f64)
        pk = np.zeros(d, dtype=np.int32) if prior_kind is None else np.ascontiguousarray(prior_kind, dtype=np.int32)
        pa = np.zeros(d) if prior_a is None else np.ascontiguousarray(prior_a, dtype=np.float64)
        pb = np.ones(d) if prior_b is None else np.ascontiguousarray(prior_b, dtype=np.float64)
        mins, maxs = np.full(d, -np.inf), np.full(d, np.inf)
        u = pk == _cabi.PRIOR_UNIFORM
        mins[u], maxs[u] = pa[u], pa[u] + pb[u]
        self.all_flat = bool(np.all(pk == _cabi.PRIOR_FLAT))
        self.prior_kind = torch.from_numpy(pk).to(self.device)
        self.prior_a, self.prior_b = torch.from_numpy(pa).to(self.device), torch.from_numpy(pb).to(self.device)
        self.mins, self.maxs = torch.from_numpy(mins).to(self.device), torch.from_numpy(maxs).to(self.device)
        crp = np.array(cr_probs if cr_probs is not None else [1 / float(nCR)] * nCR, dtype=np.float64)
        gpr = np.array(gamma_probs if gamma_probs is not None else [1 / float(gamma_levels)] * gamma_levels, dtype=np.float64)
        self.cr_probs, self.gamma_probs = torch.from_numpy(crp).to(self.device), torch.from_numpy(gpr).to(self.device)
        self.ncr_updates, self.delta_m = torch.zeros(nCR, **f64), torch.zeros(nCR, **f64)
        self.ngamma_updates, self.delta_m_gamma = torch.zeros(gamma_levels, **f64), torch.zeros(gamma_levels, **f64)
        self.gamma_table = torch.from_numpy(gamma_table(gamma_levels, DEpairs, d)).to(self.device)
        self.target_table = torch.from_numpy(device_target_table(target, self.ld)).to(self.device)
        # carried y = invC x and Q = x.y of the dense-Gaussian window kernel (include/dreamzs.h dreamzs_state)
        dense = target.kind == T.TARGET_GAUSSIAN_DENSE
        self.gauss_Y = torch.zeros((self.Nl, self.ld), **f64) if dense else None
        self.gauss_Q = torch.zeros(self.Nl, **f64) if dense else None
        # whitened form (dreamzs_state.gauss_L / gauss_U): invC = L L^T, carried u = L^T x
        self.gauss_L = self.gauss_U = self.sync_ws = None
        if dense and whitened and self.ld <= 128:
            cache = target.__dict__.setdefault('_whitening_cache', {})     # the factor depends on the target only
            if self.ld not in cache:
                Lw = whitening_factor(target.invC)
                cache[self.ld] = None if Lw is None else pack_whitening(Lw, self.ld)
            packed = cache[self.ld]
            if packed is not None:
                assert packed.size == int(self.lib.dreamzs_whiten_doubles(self.ld))
                self.gauss_L = torch.from_numpy(packed).to(self.device)
                self.gauss_U = torch.zeros((self.Nl, self.ld), **f64)
                # scratch of the persistent multi-window launches (dreamzs_state.sync_ws): abort word + one counter per window
                self.sync_ws = torch.zeros(16 + 4096 + _cabi.SYNC_GROUP_WORDS, dtype=torch.int32, device=self.device)
        self.cfg = _cabi.Config(abi_version=_cabi.ABI_VERSION, ndim=d, ld=self.ld, nchains_global=N, chain_begin=self.c0,
                                nchains_local=self.Nl, nCR=nCR, ngamma=gamma_levels, nDEpairs=DEpairs, multitry=multitry,
                                hardboundaries=int(bool(hardboundaries)), history_thin=self.thin,
                                target_kind=int(target.kind), flags=(_cabi.FLAG_ALL_FLAT if self.all_flat else 0) | (_cabi.FLAG_GENERIC_KERNEL if generic_kernel else 0)
                                | (0 if window_kernel else _cabi.FLAG_NO_WINDOW_KERNEL),
                                snooker=snooker, p_gamma_unity=p_gamma_unity, lamb=lamb, zeta=zeta, seed=int(seed) & (2 ** 64 - 1))
        # scratch of the two-stage multi-try step (dreamzs_state.draw_ws): every draw of a window, made ahead of the chains
        # (the dense-Gaussian window kernels need none)
        windowed = dense and self.all_flat and DEpairs == 1 and multitry == 1 and window_kernel and not generic_kernel
        nb = int(self.lib.dreamzs_draw_ws_bytes(C.byref(self.cfg), min(self.thin, int(draw_iters or self.thin)))) if (two_stage and not self.external and not windowed) else 0
        self.draw_ws = torch.empty(nb // 8, **f64) if 0 < nb <= 2 ** 31 else None
        ws = self.lib.dreamzs_adapt_workspace_bytes(C.byref(self.cfg))
        self.workspace = torch.zeros(max(int(ws), 8) // 8 + 1, **f64)
        self.colsum, self.colsq = torch.zeros(d, **f64), torch.zeros(d, **f64)
        self.partial = torch.zeros(2 * nCR + 2 * gamma_levels, **f64)
        self.launches = 0
        self._hook = _cabi.APPEND_HOOK(self._append_hook)
        self._reduce = _cabi.REDUCE_HOOK(self._reduce_hook)
        self._state()
        _cabi.check(self.lib.dreamzs_init_logp(C.byref(self.cfg), C.byref(self.st), self._stream()), 'dreamzs_init_logp')
        self.launches += 1
        if self.external:
            k = int(multitry)
            self._mt = k
            self._prop = torch.zeros((self.Nl, 2 * k - 1, self.ld), **f64)     # proposals, then the reference set
            self._aux = torch.zeros((self.Nl, 4 if k == 1 else 4 * k + 4), **f64)
            self._like = torch.zeros((self.Nl, 2 * k - 1), **f64)
            self._ext_error = torch.zeros(1, dtype=torch.int32, device=self.device)
            self.last_like.copy_(target.evaluate(self.X[:, :d].contiguous()))     # first-call logp, Dream.py:266-268

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _state(self):
        p = lambda t: C.c_void_p(t.data_ptr())
        self.st = _cabi.State(Z=p(self.Z), Z_capacity_rows=self.Z.shape[0], X=p(self.X), last_prior=p(self.last_prior),
                              last_like=p(self.last_like), cr_probs=p(self.cr_probs), gamma_probs=p(self.gamma_probs),
                              gamma_table=p(self.gamma_table), target_table=p(self.target_table),
                              prior_kind=p(self.prior_kind), prior_a=p(self.prior_a), prior_b=p(self.prior_b),
                              mins=p(self.mins), maxs=p(self.maxs),
                              gauss_Y=p(self.gauss_Y) if self.gauss_Y is not None else None,
                              gauss_Q=p(self.gauss_Q) if self.gauss_Q is not None else None,
                              gauss_L=p(self.gauss_L) if self.gauss_L is not None else None,
                              gauss_U=p(self.gauss_U) if self.gauss_U is not None else None,
                              sync_ws=p(self.sync_ws) if self.sync_ws is not None and self.persistent else None,
                              sync_ws_words=self.sync_ws.numel() if self.sync_ws is not None and self.persistent else 0,
                              draw_ws=p(self.draw_ws) if self.draw_ws is not None else None,
                              draw_ws_bytes=self.draw_ws.numel() * 8 if self.draw_ws is not None else 0)

    def _ensure_capacity(self, rows):
        """Make room for `rows` archive rows.  Collective when the archive is shared between ranks."""
        if self.Z is not None and self.Z.shape[0] >= rows:
            return
        old = self.Z
        if old is not None:
            # geometric growth: an astep() / run() loop that appends a few rows per call must not re-allocate and copy the
            # whole archive every time (and, sharded, pay a barrier + IPC re-mapping every time)
            rows = max(int(rows), old.shape[0] + old.shape[0] // 2 + self.N)
        if self._want_peers:
            Z = self._alloc_shared(rows)
        else:
            # rows past the current size are written (record_history) before they can be sampled: no zero-fill
            Z = torch.empty((rows, self.ld), dtype=torch.float64, device=self.device)
        if old is None:
            # seed rows: one host->device copy of the caller's array (asynchronous when it lives in pinned memory),
            # padded to the row stride on the device
            h = torch.from_numpy(np.ascontiguousarray(self._hist_host))
            hd = h.to(self.device, non_blocking=h.is_pinned())
            if self.ld == self.d:
                Z[:self.nseed] = hd
            else:
                Z[:self.nseed, :self.d] = hd
                Z[:self.nseed, self.d:] = 0
            self._hist_host = None
        else:
            Z[:self.archive_rows] = old[:self.archive_rows]
        self.Z = Z
        if self.peers is not None:
            self._finish_shared()
        if getattr(self, '_shared_stale', None) is not None:      # fell back to NCCL while growing: drop the old mappings
            torch.cuda.synchronize(self.device)
            torch.distributed.barrier(self.group)
            self._shared, stale = self._shared_stale, None
            self._shared_stale = None
            self._release_shared()
        if hasattr(self, 'st'):
            self._state()

    def _alloc_shared(self, rows):
        """Archive block other ranks can map (dreamzs_shared_alloc + handle exchange).  Falls back to a private
        archive + NCCL all-gather (for every rank alike) when a mapping cannot be made."""
        dist = torch.distributed
        lib = self.lib
        nbytes = SHARED_HEADER + rows * self.ld * 8
        key = (id(self.group), self.device.index)
        ent = _SHARED_CACHE.get(key)
        if ent is not None and self._shared is None and ent['nbytes'] >= nbytes and ent['world'] == self.world:
            # (every rank takes this branch together: sizes and call history are the same on all of them)
            del _SHARED_CACHE[key]
            block = ent['block']
            block[:SHARED_HEADER].zero_()             # append flags, error word, counters, per-group progress words
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)                  # nobody publishes into a header that is still being zeroed
            self._shared_new = dict(base=ent['base'], opened=ent['opened'], block=block, nbytes=ent['nbytes'])
            self._fill_peers(ent['base'], ent['opened'])
            cap_rows = (ent['nbytes'] - SHARED_HEADER) // (self.ld * 8)
            return block[SHARED_HEADER:SHARED_HEADER + cap_rows * self.ld * 8].view(torch.float64).view(cap_rows, self.ld)
        if ent is not None and self._shared is None:  # cached block too small (or of another world size): free it first
            del _SHARED_CACHE[key]
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)
            for q, pq in enumerate(ent['opened']):
                if q != ent['rank']:
                    lib.dreamzs_shared_close(C.c_void_p(pq))
            lib.dreamzs_shared_free(C.c_void_p(ent['base']))
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)                      # nobody is still writing into the block being replaced
        base, handle = C.c_void_p(), C.create_string_buffer(64)
        ok = lib.dreamzs_shared_alloc(nbytes, C.byref(base), handle) == _cabi.OK
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw if ok else None, group=self.group)
        opened = []
        if ok and all(h is not None for h in handles):
            for q, h in enumerate(handles):
                if q == self.rank:
                    opened.append(base.value)
                    continue
                pq = C.c_void_p()
                if lib.dreamzs_shared_open(C.create_string_buffer(h, 64), C.byref(pq)) != _cabi.OK:
                    ok = False
                    break
                opened.append(pq.value)
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:                     # some rank cannot map a peer: every rank uses NCCL
            for q, pq in enumerate(opened):
                if q != self.rank:
                    lib.dreamzs_shared_close(C.c_void_p(pq))
            if base.value:
                lib.dreamzs_shared_free(base)
            self._want_peers = False
            # from here on every rank keeps a private archive and the NCCL all-gather carries the appended rows; an
            # earlier shared block (the archive was growing) is released once _ensure_capacity has copied it
            self.peers = None
            self._shared_stale = self._shared
            self._shared = None
            return torch.empty((rows, self.ld), dtype=torch.float64, device=self.device)
        block = torch.as_tensor(_DevicePtr(base.value, nbytes), device=self.device)
        if self._shared is not None:                  # growing: the flags published so far carry over
            old_block = self._shared['block']
            block[:SHARED_HEADER].copy_(old_block[:SHARED_HEADER])
        self._shared_new = dict(base=base.value, opened=opened, block=block, nbytes=nbytes)
        self._fill_peers(base.value, opened)
        return block[SHARED_HEADER:].view(torch.float64).view(rows, self.ld)

    def _fill_peers(self, base, opened):
        pr = _cabi.Peers(world=self.world, rank=self.rank)
        for q, pq in enumerate(opened):
            pr.Z[q] = pq + SHARED_HEADER
            pr.flags[q] = pq
        pr.counter = base + 2048
        pr.error = base + 1024
        pr.gflag_stride = GFLAG_STRIDE
        self.peers = pr

    def _finish_shared(self):
        """Second half of a (re)allocation of the shared archive: every rank's block is filled, then the old
        mappings go away."""
        if not hasattr(self, '_shared_new') or self._shared_new is None:
            return
        torch.cuda.synchronize(self.device)
        torch.distributed.barrier(self.group)         # every replica holds its seed rows / old contents
        self._release_shared()
        self._shared, self._shared_new = self._shared_new, None

    def _release_shared(self):
        if self._shared is None:
            return
        for q, pq in enumerate(self._shared['opened']):
            if q != self.rank:
                self.lib.dreamzs_shared_close(C.c_void_p(pq))
        self.lib.dreamzs_shared_free(C.c_void_p(self._shared['base']))
        self._shared = None

    def close(self, cache=True):
        """Collective: give the shared archive block back (kept mapped for the next engine of this process group unless
        cache=False: then the peers' archives are unmapped and the block is freed).  No-op for private archives."""
        if self._shared is not None:
            torch.cuda.synchronize(self.device)
            torch.distributed.barrier(self.group)
            self.Z = None
            self.peers = None
            key = (id(self.group), self.device.index)
            if cache and key not in _SHARED_CACHE:
                _SHARED_CACHE[key] = dict(self._shared, group=self.group, device=self.device, rank=self.rank, world=self.world)
                self._shared = None
            else:
                self._release_shared()

    def rewind(self):
        """Collective.  Forget the rows appended so far (the archive is back to its seed rows) and restart the iteration
        counter at 0; the chains keep their positions.  The next run re-uses the same archive rows, so a benchmark can
        repeat a run of any length inside a fixed amount of memory."""
        torch.cuda.synchronize(self.device)
        if self.group is not None:
            torch.distributed.barrier(self.group)     # no peer is still storing rows / flags into this replica
        self.count, self.iter = 0, 0
        if self._shared is not None:
            self._shared['block'][:SHARED_HEADER].zero_()       # append flags, error word, appended-chains counter
            torch.cuda.synchronize(self.device)
            torch.distributed.barrier(self.group)
        if not self.external:
            # first-call branch again: last_prior / last_like (and the window kernel's carried state) from X
            _cabi.check(self.lib.dreamzs_init_logp(C.byref(self.cfg), C.byref(self.st), self._stream()), 'dreamzs_init_logp')
            self.launches += 1

    def check_peers(self):
        """Raise if a wait for a peer's append timed out (DREAMZS_PEER_TIMEOUT_NS), or if a multi-try batch of the split
        step still had no finite log-posterior after 1000 regenerated batches (Dream.py:278-289 would loop on)."""
        if self.external and int(self._ext_error.item()) != 0:
            raise _cabi.DreamzsError('every proposal of a multi-try batch had a non-finite log-posterior, 1000 regenerated '
                                     'batches included')
        if self._shared is not None:
            err = self._shared['block'][1024:1028].view(torch.int32)
            if int(err.item()) != 0:
                raise _cabi.DreamzsError('timed out waiting for a peer GPU to publish its archive append')
        if self.sync_ws is not None and self.persistent and int(self.sync_ws[0].item()) != 0:
            raise _cabi.DreamzsError('a persistent launch gave up waiting for rows appended inside it (abort word set)')

    @property
    def archive_rows(self):
        return self.nseed + self.count

    # ------------------------------------------------------------------ sampling
    def run(self, niter, trace=None, logp=None, decisions=None):
        """Run `niter` iterations for the local chains.  Returns (trace [Nl,niter,ld], logp [Nl,niter],
        decisions [Nl,niter] or None) as device tensors (stream-ordered, not synchronised)."""
        niter = int(niter)
        dev = self.device
        if trace is None:
            trace = torch.empty((self.Nl, niter, self.ld), dtype=torch.float64, device=dev)
        if logp is None:
            logp = torch.empty((self.Nl, niter), dtype=torch.float64, device=dev)
        if decisions is None and self.record_decisions:
            decisions = torch.empty((self.Nl, niter), dtype=torch.int32, device=dev)
        self._ensure_capacity(self.archive_rows + appends_in(self.iter, niter, self.thin) * self.N)
        self._advance(niter, trace, logp, decisions)
        return trace, logp, decisions

    def _advance(self, niter, trace, logp, decisions):
        """`niter` iterations into trace rows 0..niter-1 of the given buffers ([Nl, >=niter, ld] etc.).
        Burn-in iterations with adaptation run one launch each with the reduction kernels in between
        (Dream.py:364-401); everything else goes through the native window loop dreamzs_run."""
        adapting = self.adapt_crossover or self.adapt_gamma
        T_ = trace.shape[1]
        tr = _cabi.Trace(trace=trace.data_ptr(), trace_logp=logp.data_ptr(),
                         decisions=decisions.data_ptr() if decisions is not None else None, trace_iters=T_, trace_offset=0)
        stream = self._stream()
        cfg, st = C.byref(self.cfg), C.byref(self.st)
        t_first, end = self.iter, self.iter + niter
        t = self.iter
        if self.external:
            self._advance_external(niter, trace, logp, decisions, tr)
            return
        hook = self._hook if (self.world > 1 and self.peers is None) else _cabi.APPEND_HOOK()
        peers = C.byref(self.peers) if self.peers is not None else None
        adapt = None
        if adapting and t <= self.crossover_burnin:
            # burn-in adaptation runs inside the native loop (dreamzs_adapt): one launch per iteration + reductions
            p = lambda x: C.c_void_p(x.data_ptr())
            if getattr(self, '_x_entry', None) is None:
                self._x_entry = torch.empty_like(self.X)
            self._adapt_ctx = _cabi.Adapt(
                adapt_crossover=int(self.adapt_crossover), adapt_gamma=int(self.adapt_gamma), crossover_burnin=self.crossover_burnin,
                colsum=p(self.colsum), colsq=p(self.colsq), partial=p(self.partial), workspace=p(self.workspace),
                x_entry=p(self._x_entry), ncr_updates=p(self.ncr_updates), delta_m=p(self.delta_m), cr_probs=p(self.cr_probs),
                ngamma_updates=p(self.ngamma_updates), delta_m_gamma=p(self.delta_m_gamma), gamma_probs=p(self.gamma_probs),
                reduce=self._reduce if self.world > 1 else _cabi.REDUCE_HOOK(), user=None)
            adapt = C.byref(self._adapt_ctx)
        tr.trace_offset = 0
        nl, rows = C.c_int64(0), C.c_int64(0)
        rc = self.lib.dreamzs_run(cfg, st, C.byref(tr), t, end - t, self.archive_rows, self.count // self.N, peers, hook,
                                  None, adapt, stream, C.byref(nl), C.byref(rows))
        _cabi.check(rc, 'dreamzs_run')
        self.launches += int(nl.value)
        self.count = int(rows.value) - self.nseed
        self.iter = end

    def _advance_external(self, niter, trace, logp, decisions, tr):
        """Split step for a caller-evaluated likelihood: per iteration dreamzs_propose -> target.evaluate ->
        dreamzs_accept (+ the burn-in adaptation stages and the archive exchange of sharded runs)."""
        lib, cfg, st, stream = self.lib, C.byref(self.cfg), C.byref(self.st), self._stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        adapting = self.adapt_crossover or self.adapt_gamma
        t_first = self.iter
        if adapting and t_first <= self.crossover_burnin:
            x_entry = self.X.clone()
        for t in range(t_first, t_first + niter):
            k, d = self._mt, self.d
            _cabi.check(lib.dreamzs_propose(cfg, st, t, self.archive_rows, p(self._prop), p(self._aux), stream), 'dreamzs_propose')
            self._like[:, :k] = self.target.evaluate(self._prop[:, :k, :d].reshape(-1, d).contiguous()).view(self.Nl, k)
            if k > 1:      # multi-try: pick a proposal, then the reference set around it (Dream.py:291-303)
                for _ in range(1000):    # Dream.py:278-289: chains without any finite proposal draw the next batch
                    self._ext_error.zero_()
                    _cabi.check(lib.dreamzs_repropose(cfg, st, t, self.archive_rows, p(self._prop), p(self._aux), p(self._like),
                                                      p(self._ext_error), stream), 'dreamzs_repropose')
                    self.launches += 1
                    if int(self._ext_error.item()) == 0:
                        break
                    self._like[:, :k] = self.target.evaluate(self._prop[:, :k, :d].reshape(-1, d).contiguous()).view(self.Nl, k)
                _cabi.check(lib.dreamzs_select(cfg, st, t, self.archive_rows, p(self._prop), p(self._aux), p(self._like),
                                               p(self._ext_error), stream), 'dreamzs_select')
                self._like[:, k:] = self.target.evaluate(self._prop[:, k:, :d].reshape(-1, d).contiguous()).view(self.Nl, k - 1)
                self.launches += 1
            tr.trace_offset = t - t_first
            _cabi.check(lib.dreamzs_accept(cfg, st, C.byref(tr), t, self.archive_rows, p(self._prop), p(self._aux), p(self._like), stream),
                        'dreamzs_accept')
            self.launches += 2
            if adapting and ((10 < t < self.crossover_burnin) or t == self.crossover_burnin):
                trow, T_ = t - t_first, trace.shape[1]
                x_old, ld_old = (p(x_entry), self.ld) if trow == 0 else (C.c_void_p(trace.data_ptr() + (trow - 1) * self.ld * 8), T_ * self.ld)
                self._adapt_stage(t, x_old, ld_old, C.c_void_p(decisions.data_ptr() + trow * 4), T_)
            if t % self.thin == 0:       # record_history for every chain; sharded: gather the other ranks' rows
                M = self.archive_rows
                allgather_rows(self.Z[M:M + self.N], self.c0, self.Nl, self.group)
                self.count += self.N
        self.iter = t_first + niter

    def astep(self, q0=None, T=1., last_loglike=None, last_logprior=None):
        """The step-operator contract of the reference, `Dream.astep(q0, T, last_loglike, last_logprior) ->
        (q_new, log_prior, log_like)` (pydream/Dream.py:193, 422; called by name in core.py:97, 114 and 232), for every
        local chain at once: one iteration (appends to the archive when iter % history_thin == 0, burn-in
        adaptation included).  q0 [N_local, ndim] replaces the current positions (None: continue from them);
        last_loglike / last_logprior [N_local] replace the stored values as in Dream.py:240-243 (None with a new q0:
        they are re-evaluated, the first-call branch Dream.py:266-268); T is a scalar or one temperature per local
        chain.  Returns device tensors (q_new [N_local, ndim], log_prior [N_local], log_like [N_local])."""
        f64 = dict(dtype=torch.float64, device=self.device)
        lib, cfg, stream = self.lib, C.byref(self.cfg), self._stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        if self.external:
            raise NotImplementedError('astep needs an analytic target; caller-evaluated likelihoods go through run()')
        if q0 is not None:
            self.X[:, :self.d] = torch.as_tensor(q0, **f64).reshape(self.Nl, self.d)
            # first-call branch (also re-derives the window kernel's carried y = invC x, Q = x.y)
            _cabi.check(lib.dreamzs_init_logp(cfg, C.byref(self.st), stream), 'dreamzs_init_logp')
            self.launches += 1
        if last_loglike is not None:
            self.last_like.copy_(torch.as_tensor(last_loglike, **f64).reshape(self.Nl))
            self.last_prior.copy_(torch.as_tensor(last_logprior, **f64).reshape(self.Nl))
        trace = torch.empty((self.Nl, 1, self.ld), **f64)
        logp = torch.empty((self.Nl, 1), **f64)
        dec = torch.empty((self.Nl, 1), dtype=torch.int32, device=self.device)
        t = self.iter
        self._ensure_capacity(self.archive_rows + appends_in(t, 1, self.thin) * self.N)
        Tt = torch.as_tensor(T, **f64)
        if Tt.numel() == 1 and float(Tt) == 1.0:
            self._advance(1, trace, logp, dec)          # fused kernels, sharded archives, adaptation
        else:
            if self.world > 1:
                raise NotImplementedError('tempered steps run on one GPU')
            self.temperature = Tt.expand(self.Nl).contiguous() if Tt.numel() == 1 else Tt.reshape(self.Nl).contiguous()
            adapting = (self.adapt_crossover or self.adapt_gamma) and ((10 < t < self.crossover_burnin) or t == self.crossover_burnin)
            x_entry = self.X.clone() if adapting else None
            tr = _cabi.Trace(trace=trace.data_ptr(), trace_logp=logp.data_ptr(), decisions=dec.data_ptr(), trace_iters=1, trace_offset=0)
            _cabi.check(lib.dreamzs_step_tempered(cfg, C.byref(self.st), C.byref(tr), t, self.archive_rows, p(self.temperature), stream),
                        'dreamzs_step_tempered')
            self.launches += 1
            if adapting:
                self._adapt_stage(t, p(x_entry), self.ld, p(dec), 1)
            if t % self.thin == 0:
                self.count += self.N
            self.iter = t + 1
            if self.gauss_Y is not None:    # keep the window kernel's carried state valid for an untempered continuation
                _cabi.check(lib.dreamzs_init_logp(cfg, C.byref(self.st), stream), 'dreamzs_init_logp')
                self.launches += 1
        self.last_decisions = dec[:, 0]
        return trace[:, 0, :self.d], self.last_prior.clone(), self.last_like.clone()

    def _adapt_stage(self, t, x_old, ld_old, dec, dec_stride):
        """One sweep of estimate_crossover_probabilities / estimate_gamma_level_probs (Dream.py:451-540) after
        iteration t: x_old / ld_old = the states before the iteration (pointer, chain stride in doubles), dec /
        dec_stride = this iteration's decision words."""
        lib, cfg, stream = self.lib, C.byref(self.cfg), self._stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        _cabi.check(lib.dreamzs_adapt_colsum(cfg, p(self.X), p(self.colsum), p(self.workspace), stream), 'dreamzs_adapt_colsum')
        allreduce_sum(self.colsum, self.group)
        _cabi.check(lib.dreamzs_adapt_colsq(cfg, p(self.X), p(self.colsum), p(self.colsq), p(self.workspace), stream), 'dreamzs_adapt_colsq')
        allreduce_sum(self.colsq, self.group)
        _cabi.check(lib.dreamzs_adapt_jumps(cfg, p(self.X), x_old, ld_old, dec, dec_stride, p(self.colsq), int(t == self.crossover_burnin),
                                            int(self.adapt_crossover), int(self.adapt_gamma), p(self.partial), p(self.workspace), stream),
                    'dreamzs_adapt_jumps')
        allreduce_sum(self.partial, self.group)
        _cabi.check(lib.dreamzs_adapt_finish(cfg, p(self.partial), int(self.adapt_crossover), int(self.adapt_gamma),
                                             p(self.ncr_updates), p(self.delta_m), p(self.cr_probs), p(self.ngamma_updates),
                                             p(self.delta_m_gamma), p(self.gamma_probs), stream), 'dreamzs_adapt_finish')
        self.launches += 7

    def run_tempered(self, niter, temperature=None):
        """Parallel tempering, the loop of _sample_dream_pt (pydream/core.py:131-236): per iteration one astep of every
        chain at its temperature (dreamzs_step_tempered; trace row 2t), the burn-in adaptation sweep if any, then the
        proposed exchange between two chains (dreamzs_pt_swap; trace row 2t+1).  `temperature` defaults to the
        reference's ladder T[i] = 0.001 ** (i / nchains).  Returns device tensors (trace [N, 2 niter, ld],
        logp [N, 2 niter] = T like + prior, decisions [N, 2 niter], swaps [niter, 4] = first chain, second chain,
        accepted, alpha).  All chains must live on this GPU (the exchanged pair may be any two chains)."""
        if self.world > 1:
            raise NotImplementedError('parallel tempering exchanges states between arbitrary chains: run it on one GPU')
        if self.external:
            raise NotImplementedError('parallel tempering needs an analytic target (the split step carries no temperature)')
        niter = int(niter)
        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)
        if temperature is None:
            temperature = temperature_ladder(self.N)
        self.temperature = torch.as_tensor(np.ascontiguousarray(temperature, dtype=np.float64)).to(dev)
        if self.temperature.numel() != self.N:
            raise ValueError('one temperature per chain')
        trace = torch.empty((self.Nl, 2 * niter, self.ld), **f64)
        logp = torch.empty((self.Nl, 2 * niter), **f64)
        decisions = torch.zeros((self.Nl, 2 * niter), dtype=torch.int32, device=dev)
        swaps = torch.zeros((niter, 8), **f64)
        self._ensure_capacity(self.archive_rows + appends_in(self.iter, niter, self.thin) * self.N)
        lib, cfg, st, stream = self.lib, C.byref(self.cfg), C.byref(self.st), self._stream()
        p = lambda x: C.c_void_p(x.data_ptr())
        tr = _cabi.Trace(trace=trace.data_ptr(), trace_logp=logp.data_ptr(), decisions=decisions.data_ptr(),
                         trace_iters=2 * niter, trace_offset=0)
        adapting = self.adapt_crossover or self.adapt_gamma
        t_first = self.iter
        x_entry = self.X.clone() if adapting and t_first <= self.crossover_burnin else None
        for t in range(t_first, t_first + niter):
            row = 2 * (t - t_first)
            tr.trace_offset = row
            _cabi.check(lib.dreamzs_step_tempered(cfg, st, C.byref(tr), t, self.archive_rows, p(self.temperature), stream),
                        'dreamzs_step_tempered')
            self.launches += 1
            if adapting and ((10 < t < self.crossover_burnin) or t == self.crossover_burnin):
                # q0 of this astep is the state after the previous iteration's exchange (trace row 2t-1)
                x_old, ld_old = (p(x_entry), self.ld) if row == 0 else (C.c_void_p(trace.data_ptr() + (row - 1) * self.ld * 8), 2 * niter * self.ld)
                self._adapt_stage(t, x_old, ld_old, C.c_void_p(decisions.data_ptr() + row * 4), 2 * niter)
            if t % self.thin == 0:
                self.count += self.N
            tr.trace_offset = row + 1
            _cabi.check(lib.dreamzs_pt_swap(cfg, st, C.byref(tr), t, p(self.temperature), p(swaps[t - t_first]), stream), 'dreamzs_pt_swap')
            self.launches += 2
        self.iter = t_first + niter
        if self.gauss_Y is not None:
            # the window kernel's carried y = invC x / Q = x.y went stale (the tempered step does not maintain them):
            # re-derive them (and, identically, last_prior / last_like) in case the engine continues untempered
            _cabi.check(lib.dreamzs_init_logp(cfg, st, stream), 'dreamzs_init_logp')
            self.launches += 1
        return trace, logp, decisions, swaps[:, :4]

    def run_to_host(self, niter, out_params, out_logp, chunk_iters=256, on_chunk=None):
        """Run `niter` iterations and stream the samples to host memory while sampling continues.
        out_params: pinned host tensor [Nl, niter, d]; out_logp: pinned host tensor [Nl, niter] (or [Nl, niter, 1]).
        The run is cut into chunks of `chunk_iters` iterations written to two alternating device buffers; each
        finished chunk leaves on a copy stream (pitched device->host copies through the C ABI) while the next
        one is being sampled.  `on_chunk(decisions_chunk, t0)` (optional) sees each chunk's decision words."""
        niter, d, ld, Nl = int(niter), self.d, self.ld, self.Nl
        dev = self.device
        Tc = max(1, min(int(chunk_iters), niter))
        self._ensure_capacity(self.archive_rows + appends_in(self.iter, niter, self.thin) * self.N)
        compute = torch.cuda.current_stream(dev)
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy = self._copy_stream
        need_dec = self.record_decisions
        bufs = []
        for _ in range(2 if niter > Tc else 1):
            bufs.append(dict(trace=torch.empty((Nl, Tc, ld), dtype=torch.float64, device=dev),
                             logp=torch.empty((Nl, Tc), dtype=torch.float64, device=dev),
                             dec=torch.empty((Nl, Tc), dtype=torch.int32, device=dev) if need_dec else None,
                             packed=torch.empty((Nl, Tc, d), dtype=torch.float64, device=dev) if ld != d else None,
                             filled=torch.cuda.Event(), drained=None))
        hp, hl = out_params.data_ptr(), out_logp.data_ptr()
        p = lambda t: C.c_void_p(t.data_ptr())
        t0, k = 0, 0
        while t0 < niter:
            n = min(Tc, niter - t0)
            b = bufs[k % len(bufs)]
            if b['drained'] is not None:
                compute.wait_event(b['drained'])          # the copy of the chunk that used this buffer has finished
            self._advance(n, b['trace'], b['logp'], b['dec'])
            src, spitch = b['trace'], Tc * ld * 8
            if ld != d:                                   # strip the row padding on the device
                b['packed'][:, :n].copy_(b['trace'][:, :n, :d])
                src, spitch = b['packed'], Tc * d * 8
            if on_chunk is not None and b['dec'] is not None:
                on_chunk(b['dec'][:, :n], t0)
            b['filled'].record(compute)
            copy.wait_event(b['filled'])
            cs = C.c_void_p(copy.cuda_stream)
            _cabi.check(self.lib.dreamzs_copy_d2h_2d(C.c_void_p(hp + t0 * d * 8), niter * d * 8, p(src), spitch, n * d * 8, Nl, cs),
                        'dreamzs_copy_d2h_2d')
            _cabi.check(self.lib.dreamzs_copy_d2h_2d(C.c_void_p(hl + t0 * 8), niter * 8, p(b['logp']), Tc * 8, n * 8, Nl, cs),
                        'dreamzs_copy_d2h_2d')
            b['drained'] = torch.cuda.Event()
            b['drained'].record(copy)
            t0 += n
            k += 1
        compute.wait_stream(copy)                          # results are complete once the caller syncs its stream

    def _append_hook(self, user, first_row, nrows):
        """dreamzs_append_hook: all-gather the rows the other shards appended (enqueued on the current stream)."""
        try:
            allgather_rows(self.Z[first_row:first_row + nrows], self.c0, self.Nl, self.group)
            return 0
        except Exception:      # never let an exception cross the C ABI
            import traceback
            traceback.print_exc()
            return _cabi.E_LAUNCH

    def _reduce_hook(self, user, ptr, count):
        """dreamzs_reduce_hook: sum one of the adaptation buffers over the ranks (enqueued on the current stream)."""
        try:
            for t in (self.colsum, self.colsq, self.partial):
                if t.data_ptr() == ptr:
                    allreduce_sum(t, self.group)
                    return 0
            return _cabi.E_BADARG
        except Exception:      # never let an exception cross the C ABI
            import traceback
            traceback.print_exc()
            return _cabi.E_LAUNCH

    # ------------------------------------------------------------------ diagnostics / export
    def gelman_rubin(self, trace):
        """Gelman_Rubin (pydream/convergence.py:3-20) of a device trace [Nl, T, ld] -> Rhat[d] (device)."""
        return gelman_rubin_device(trace, self.d, self.group)

    def history_flat(self):
        """The archive in the reference's on-disk layout: flat float64, rows of ndim (Dream.py:919-945)."""
        n = self.archive_rows
        return self.Z[:n, :self.d].contiguous().reshape(-1).cpu().numpy()


def gelman_rubin_device(trace, ndim, group=None):
    lib = _cabi.load()
    Nl, T_, ld = trace.shape
    dev = trace.device
    mean = torch.empty((Nl, ndim), dtype=torch.float64, device=dev)
    var = torch.empty((Nl, ndim), dtype=torch.float64, device=dev)
    s = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    _cabi.check(lib.dreamzs_gr_chain_stats(p(trace), Nl, T_, T_ // 2, ndim, ld, p(mean), p(var), s), 'dreamzs_gr_chain_stats')
    if group is not None and torch.distributed.get_world_size(group) > 1:
        W = torch.distributed.get_world_size(group)
        gm = torch.empty((W * Nl, ndim), dtype=torch.float64, device=dev)
        gv = torch.empty((W * Nl, ndim), dtype=torch.float64, device=dev)
        torch.distributed.all_gather_into_tensor(gm.view(-1), mean.view(-1), group=group)
        torch.distributed.all_gather_into_tensor(gv.view(-1), var.view(-1), group=group)
        mean, var = gm, gv
    rhat = torch.empty(ndim, dtype=torch.float64, device=dev)
    _cabi.check(lib.dreamzs_gr_finish(p(mean), p(var), mean.shape[0], T_, ndim, p(rhat), s), 'dreamzs_gr_finish')
    return rhat
