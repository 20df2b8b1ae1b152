"""ctypes binding of libdreamzs.so (include/dreamzs.h).  No torch types cross this boundary:
device pointers are passed as integers (``tensor.data_ptr()``), streams as ``cudaStream_t``.

The product path has no CPU fallback: if the library is missing or cannot be loaded,
`load()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DREAMZS_LIB') or os.path.join(HERE, 'libdreamzs.so')   # DREAMZS_LIB: A/B testing of builds

ABI_VERSION = 4
OK, E_BADARG, E_LAUNCH, E_UNSUPPORTED = 0, -1, -2, -3
MAX_NCR, MAX_NGAMMA, MAX_DEPAIRS, MAX_MULTITRY, MAX_NDIM = 16, 8, 8, 16, 1024
FLAG_ALL_FLAT = 1
FLAG_GENERIC_KERNEL = 2
FLAG_NO_WINDOW_KERNEL = 4
PRIOR_FLAT, PRIOR_NORMAL, PRIOR_UNIFORM = 0, 1, 2
DECISION_SWAPPED = 1 << 20

EXPORTS = ['dreamzs_abi_version', 'dreamzs_init_logp', 'dreamzs_step', 'dreamzs_run', 'dreamzs_copy_d2h_2d',
           'dreamzs_propose', 'dreamzs_select', 'dreamzs_accept', 'dreamzs_step_tempered', 'dreamzs_pt_swap',
           'dreamzs_shared_alloc', 'dreamzs_shared_open', 'dreamzs_shared_close', 'dreamzs_shared_free', 'dreamzs_adapt_workspace_bytes',
           'dreamzs_adapt_colsum', 'dreamzs_adapt_colsq', 'dreamzs_adapt_jumps', 'dreamzs_adapt_finish',
           'dreamzs_gr_chain_stats', 'dreamzs_gr_finish', 'dreamzs_whiten_doubles', 'dreamzs_rng_normals',
           'dreamzs_debug_set_phase_buffer', 'dreamzs_draw_ws_bytes', 'dreamzs_repropose']


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
                ('draw_ws', C.c_void_p), ('draw_ws_bytes', C.c_int64)]


class Trace(C.Structure):
    _fields_ = [('trace', C.c_void_p), ('trace_logp', C.c_void_p), ('decisions', C.c_void_p),
                ('trace_iters', C.c_int64), ('trace_offset', C.c_int64)]


MAX_PEERS = 8
GFLAG_OFFSET = 512          # uint64 words from a rank's append flags to its per-group flag rows (include/dreamzs.h)
SYNC_GROUP_WORDS = 4096


class Peers(C.Structure):
    _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('Z', C.c_void_p * MAX_PEERS), ('flags', C.c_void_p * MAX_PEERS),
                ('counter', C.c_void_p), ('error', C.c_void_p), ('gflag_stride', C.c_int32), ('reserved', C.c_int32)]


APPEND_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_int64)
REDUCE_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)


class Adapt(C.Structure):
    _fields_ = [('adapt_crossover', C.c_int32), ('adapt_gamma', C.c_int32), ('crossover_burnin', C.c_int64),
                ('colsum', C.c_void_p), ('colsq', C.c_void_p), ('partial', C.c_void_p), ('workspace', C.c_void_p),
                ('x_entry', C.c_void_p), ('ncr_updates', C.c_void_p), ('delta_m', C.c_void_p), ('cr_probs', C.c_void_p),
                ('ngamma_updates', C.c_void_p), ('delta_m_gamma', C.c_void_p), ('gamma_probs', C.c_void_p),
                ('reduce', REDUCE_HOOK), ('user', C.c_void_p)]


class DreamzsError(RuntimeError):
    pass


_lib = None


def load():
    """Load libdreamzs.so (built in-tree by pydream_b200.build).  Raises if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DreamzsError('libdreamzs.so not found at %s: build it with `python -m pydream_b200.build` '
                           '(there is no CPU fallback for the MT-DREAM(ZS) step path)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    cfgp, stp, trp = C.POINTER(Config), C.POINTER(State), C.POINTER(Trace)
    sig = {
        'dreamzs_abi_version': (C.c_int, []),
        'dreamzs_init_logp': (C.c_int, [cfgp, stp, vp]),
        'dreamzs_step': (C.c_int, [cfgp, stp, trp, i64, i32, i64, vp]),
        'dreamzs_run': (C.c_int, [cfgp, stp, trp, i64, i64, i64, i64, C.POINTER(Peers), APPEND_HOOK, vp, C.POINTER(Adapt), vp,
                                  C.POINTER(i64), C.POINTER(i64)]),
        'dreamzs_shared_alloc': (C.c_int, [i64, C.POINTER(vp), vp]),
        'dreamzs_shared_open': (C.c_int, [vp, C.POINTER(vp)]),
        'dreamzs_shared_close': (C.c_int, [vp]),
        'dreamzs_shared_free': (C.c_int, [vp]),
        'dreamzs_copy_d2h_2d': (C.c_int, [vp, i64, vp, i64, i64, i64, vp]),
        'dreamzs_propose': (C.c_int, [cfgp, stp, i64, i64, vp, vp, vp]),
        'dreamzs_select': (C.c_int, [cfgp, stp, i64, i64, vp, vp, vp, vp, vp]),
        'dreamzs_repropose': (C.c_int, [cfgp, stp, i64, i64, vp, vp, vp, vp, vp]),
        'dreamzs_accept': (C.c_int, [cfgp, stp, trp, i64, i64, vp, vp, vp, vp]),
        'dreamzs_step_tempered': (C.c_int, [cfgp, stp, trp, i64, i64, vp, vp]),
        'dreamzs_pt_swap': (C.c_int, [cfgp, stp, trp, i64, vp, vp, vp]),
        'dreamzs_adapt_workspace_bytes': (i64, [cfgp]),
        'dreamzs_adapt_colsum': (C.c_int, [cfgp, vp, vp, vp, vp]),
        'dreamzs_adapt_colsq': (C.c_int, [cfgp, vp, vp, vp, vp, vp]),
        'dreamzs_adapt_jumps': (C.c_int, [cfgp, vp, vp, i64, vp, i64, vp, i32, i32, i32, vp, vp, vp]),
        'dreamzs_adapt_finish': (C.c_int, [cfgp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]),
        'dreamzs_gr_chain_stats': (C.c_int, [vp, i64, i64, i64, i32, i64, vp, vp, vp]),
        'dreamzs_gr_finish': (C.c_int, [vp, vp, i64, i64, i32, vp, vp]),
        'dreamzs_whiten_doubles': (i64, [i32]),
        'dreamzs_rng_normals': (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, i32, vp, vp]),
        'dreamzs_debug_set_phase_buffer': (None, [vp]),
        'dreamzs_draw_ws_bytes': (i64, [cfgp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.dreamzs_abi_version() != ABI_VERSION:
        raise DreamzsError('libdreamzs.so ABI version %d != %d' % (lib.dreamzs_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


_ERR = {E_BADARG: 'bad argument', E_LAUNCH: 'CUDA launch failure', E_UNSUPPORTED: 'unsupported option combination'}


def check(rc, what):
    if rc != OK:
        raise DreamzsError('%s failed: %s (%d)' % (what, _ERR.get(rc, 'unknown error'), rc))
