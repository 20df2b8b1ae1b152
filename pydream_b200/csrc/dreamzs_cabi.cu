// extern "C" entry points of libdreamzs.so for the fused step (include/dreamzs.h) and the
// dispatch over the <G, R> kernel instantiations (one object file each, dreamzs_step_inst.cu).
#include <stdlib.h>
#include <string.h>
#include "dreamzs_step_params.cuh"
#include "dreamzs_common.cuh"

using namespace dreamzs;

#define DZ_DECL(g, r) int dreamzs_launch_step_##g##_##r(const dreamzs::StepParams &, int, size_t, cudaStream_t); \
  int dreamzs_launch_st2_##g##_##r(const dreamzs::StepParams &, int, size_t, cudaStream_t);
DZ_DECL(4, 1) DZ_DECL(8, 1) DZ_DECL(16, 1) DZ_DECL(32, 1) DZ_DECL(32, 2) DZ_DECL(32, 4) DZ_DECL(32, 8)
#undef DZ_DECL
#define DZ_DECL_MTP(g, r) int dreamzs_launch_mtp_##g##_##r(const dreamzs::StepParams &, cudaStream_t);
DZ_DECL_MTP(2, 1) DZ_DECL_MTP(2, 2) DZ_DECL_MTP(4, 1) DZ_DECL_MTP(4, 2) DZ_DECL_MTP(8, 1) DZ_DECL_MTP(8, 2)
#undef DZ_DECL_MTP
int dreamzs_launch_gauss_7(const dreamzs::StepParams &, cudaStream_t);
int dreamzs_launch_gauss_8(const dreamzs::StepParams &, cudaStream_t);
size_t dreamzs_launch_gauss_smem_bytes(const dreamzs_config &cfg, int TC);
int dreamzs_launch_gwin_7(dreamzs::StepParams &, cudaStream_t);
int dreamzs_launch_gwin_8(dreamzs::StepParams &, cudaStream_t);
int dreamzs_gwin_usable(const dreamzs_config &cfg, int TC);
int dreamzs_launch_gauss_refresh(const dreamzs::StepParams &, cudaStream_t);
int dreamzs_wwin_usable(const dreamzs_config &cfg, int sms, int niter);
int dreamzs_launch_wwin(dreamzs::StepParams &, int sms, cudaStream_t);
int dreamzs_launch_whiten(const dreamzs::StepParams &, cudaStream_t);
namespace dreamzs { int wwin_ntiles_host(int ld); }

static long long *g_phase_buffer = nullptr;
// profiling aid (include/dreamzs.h): CTA 0 of the window kernels writes clock64() stamps of its phases
extern "C" void dreamzs_debug_set_phase_buffer(void *device_ptr) { g_phase_buffer = (long long *)device_ptr; }

// SM count of the CURRENT device (cached per device ordinal: a process may drive several GPUs)
static int sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 148; }
  int &n = cache[dev & 63];
  if (n == 0 && (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)) n = 148;
  return n;
}

static int table_doubles_of(const dreamzs_config *cfg) {
  const int d = cfg->ndim;
  switch (cfg->target_kind) {
    case DREAMZS_TARGET_CONSTANT: case DREAMZS_TARGET_SUMSHIFT: return 1;
    case DREAMZS_TARGET_GAUSSIAN_DENSE: return 2 + cfg->ld * d;   // [log_F, 0, invC^T rows padded to ld]
    case DREAMZS_TARGET_MIXTURE: return 2 + 2 * d;
    case DREAMZS_TARGET_BANANA: return 2;
    case DREAMZS_TARGET_EXTERNAL: return 1;
    default: return -1;
  }
}

static int check_cfg(const dreamzs_config *cfg, const dreamzs_state *st) {
  if (!cfg || !st || cfg->abi_version != DREAMZS_ABI_VERSION) return DREAMZS_E_BADARG;
  if (cfg->ndim < 1 || cfg->ndim > DREAMZS_MAX_NDIM || cfg->ld < cfg->ndim || (cfg->ld & 3)) return DREAMZS_E_BADARG;
  if (cfg->nchains_local < 0 || cfg->chain_begin < 0 || cfg->chain_begin + cfg->nchains_local > cfg->nchains_global) return DREAMZS_E_BADARG;
  if (cfg->nCR < 1 || cfg->nCR > DREAMZS_MAX_NCR || cfg->ngamma < 1 || cfg->ngamma > DREAMZS_MAX_NGAMMA) return DREAMZS_E_BADARG;
  if (cfg->nDEpairs < 1 || cfg->nDEpairs > DREAMZS_MAX_DEPAIRS) return DREAMZS_E_BADARG;
  if (cfg->multitry < 1 || 2 * cfg->multitry > DREAMZS_MAX_MULTITRY) return DREAMZS_E_BADARG;
  if (cfg->multitry == 2) return DREAMZS_E_UNSUPPORTED;   // broken in the reference too (Dream.py:867-868)
  if (cfg->history_thin < 1) return DREAMZS_E_BADARG;
  if (!st->Z || !st->X || !st->last_prior || !st->last_like || !st->cr_probs || !st->gamma_probs || !st->gamma_table ||
      !st->target_table || !st->prior_kind || !st->prior_a || !st->prior_b || !st->mins || !st->maxs)
    return DREAMZS_E_BADARG;
  return DREAMZS_OK;
}

// dense Gaussian, flat priors, one DE pair, no multi-try, whitening factor given: whitened window kernel
static bool wwin_eligible(const StepParams &P) {
  const dreamzs_config &cfg = P.cfg;
  return cfg.target_kind == DREAMZS_TARGET_GAUSSIAN_DENSE && cfg.multitry == 1 && cfg.nDEpairs == 1 && P.all_flat &&
         P.st.gauss_L && P.st.gauss_U && !P.temperature && !P.ext_phase && !(cfg.flags & DREAMZS_FLAG_GENERIC_KERNEL) &&
         !(cfg.flags & DREAMZS_FLAG_NO_WINDOW_KERNEL) && cfg.ld <= 128 &&
         dreamzs_wwin_usable(cfg, sm_count(), cfg.history_thin < 16 ? cfg.history_thin : 16);
}

static bool gwin_eligible(const StepParams &P) {
  const dreamzs_config &cfg = P.cfg;
  const int chunks = cfg.ld / 4;
  return cfg.target_kind == DREAMZS_TARGET_GAUSSIAN_DENSE && cfg.multitry == 1 && cfg.nDEpairs == 1 && P.all_flat &&
         chunks > 16 && chunks <= 32 && P.st.gauss_Y && P.st.gauss_Q && !(cfg.flags & DREAMZS_FLAG_GENERIC_KERNEL) &&
         !(cfg.flags & DREAMZS_FLAG_NO_WINDOW_KERNEL) && dreamzs_gwin_usable(cfg, 8);
}

// multi-try with the k points of a batch side by side: k lane-groups of G lanes per chain, R chunks per lane (mtp_layout)
static bool mtp_eligible(const StepParams &P) {
  const dreamzs_config &cfg = P.cfg;
  int G = 0, R = 0;
  return cfg.multitry > 1 && !P.ext_phase && !P.init_only && !(cfg.flags & DREAMZS_FLAG_GENERIC_KERNEL) &&
         mtp_layout(cfg.ld, cfg.multitry, G, R);
}

// two-stage single-try step (dreamzs_st2_kernel.cuh): bytes of one iteration's records, and how many iterations' worth to
// keep.  Measured at C4 (26.5 MB per iteration): sub-spans of 2 iterations, whose records stay in L2, 209 M chain-steps/s;
// the whole window of 10 in one draw + chain pair (265 MB through HBM) 264 M -- launches cost more than the traffic
static int64_t st2_iter_bytes(const dreamzs_config &cfg) { return (int64_t)cfg.nchains_local * (4 + 2 * cfg.ld) * (int64_t)sizeof(double); }
static int64_t st2_budget() {      // DREAMZS_ST2_BUDGET_MB: experiments
  static const int64_t v = [] { const char *e = getenv("DREAMZS_ST2_BUDGET_MB"); return (int64_t)(e ? atoi(e) : 1024) << 20; }();
  return v;
}

extern "C" int64_t dreamzs_draw_ws_bytes(const dreamzs_config *cfg, int32_t niter) {
  if (!cfg || niter < 1 || cfg->nchains_local < 1) return 0;
  StepParams P{};
  P.cfg = *cfg;
  if (cfg->multitry == 1) {
    if (cfg->target_kind == DREAMZS_TARGET_EXTERNAL || (cfg->flags & DREAMZS_FLAG_GENERIC_KERNEL)) return 0;
    const int64_t per = st2_iter_bytes(*cfg);
    int64_t nb = st2_budget() / per;
    if (nb < 1) nb = 1;
    if (nb > niter) nb = niter;
    return nb * per;
  }
  if (!mtp_eligible(P)) return 0;
  return (int64_t)cfg->nchains_local * niter * (8 + (2 * cfg->multitry - 1) * 2 * cfg->ld) * (int64_t)sizeof(double);
}

static int dispatch(StepParams &P, cudaStream_t stream) {
  const dreamzs_config &cfg = P.cfg;
  const int chunks = cfg.ld / 4;
  P.table_doubles = table_doubles_of(&cfg);
  if (P.table_doubles < 0) return DREAMZS_E_UNSUPPORTED;
  P.nslots = cfg.multitry == 1 ? 1 : (P.ext_phase ? 2 * cfg.multitry - 1 : cfg.multitry + 1);
  if (wwin_eligible(P)) {
    if (P.init_only) return DREAMZS_OK;   // handled by the caller (generic evaluation + dreamzs_launch_whiten)
    P.dbg = g_phase_buffer;
    const int rc = dreamzs_launch_wwin(P, sm_count(), stream);
    if (rc != DREAMZS_E_UNSUPPORTED || P.ww_sync) return rc;
  }
  if (P.ww_sync) return DREAMZS_E_UNSUPPORTED;   // a multi-window span needs the whitened window kernel
  // dense Gaussian, flat priors, one DE pair, no multi-try, carried y = invC x: window kernel (dreamzs_gwin_kernel.cuh)
  if (gwin_eligible(P)) {
    if (P.init_only) return DREAMZS_OK;   // handled by the caller (generic evaluation + dreamzs_launch_gauss_refresh)
    const int64_t t = P.iter_begin, thin = cfg.history_thin;
    P.gw_refresh = (t == 0 || ((t - 1) % thin == 0 && ((t - 1) / thin) % DREAMZS_GAUSS_REFRESH_WINDOWS == 0)) ? 1 : 0;
    P.dbg = g_phase_buffer;
    const int tc = (cfg.nchains_local + 6) / 7 <= sm_count() ? 7 : 8;
    return tc == 7 ? dreamzs_launch_gwin_7(P, stream) : dreamzs_launch_gwin_8(P, stream);
  }
  // dense Gaussian, no multi-try, one warp-wide chunk row: CTA-synchronous kernel (dreamzs_gauss_kernel.cuh)
  if (!P.init_only && cfg.target_kind == DREAMZS_TARGET_GAUSSIAN_DENSE && cfg.multitry == 1 && chunks > 16 && chunks <= 32 &&
      !(cfg.flags & DREAMZS_FLAG_GENERIC_KERNEL) && dreamzs_launch_gauss_smem_bytes(cfg, 8) <= 227 * 1024) {
    const int tc = (cfg.nchains_local + 6) / 7 <= sm_count() ? 7 : 8;
    return tc == 7 ? dreamzs_launch_gauss_7(P, stream) : dreamzs_launch_gauss_8(P, stream);
  }
  const int threads = 128;
  int G = 32, R = 1;
  if (chunks <= 4) G = 4; else if (chunks <= 8) G = 8; else if (chunks <= 16) G = 16;
  else { R = (chunks + 31) / 32; if (R > 2 && R <= 4) R = 4; else if (R > 4) R = 8; }
  // multi-try iteration with the points of a batch side by side, a warp per chain (dreamzs_mtp_kernel.cuh): two kernels
  // (draws of the window, then the chains) when the caller gave scratch for the draws, else fused
  if (mtp_eligible(P)) {
    int g = 0, r = 0;
    mtp_layout(cfg.ld, cfg.multitry, g, r);
    if (P.st.draw_ws && P.st.draw_ws_bytes < dreamzs_draw_ws_bytes(&cfg, P.niter)) P.st.draw_ws = nullptr;
#define DZ_CASE(a, b) if (g == a && r == b) return dreamzs_launch_mtp_##a##_##b(P, stream);
    DZ_CASE(2, 1) DZ_CASE(2, 2) DZ_CASE(4, 1) DZ_CASE(4, 2) DZ_CASE(8, 1) DZ_CASE(8, 2)
#undef DZ_CASE
  }
  // single try with scratch for the draws: draw kernel + chain kernel per sub-span of the window (dreamzs_st2_kernel.cuh)
  if (cfg.multitry == 1 && !P.ext_phase && !P.init_only && P.st.draw_ws && P.st.draw_ws_bytes >= st2_iter_bytes(cfg) &&
      cfg.target_kind != DREAMZS_TARGET_EXTERNAL && !(cfg.flags & DREAMZS_FLAG_GENERIC_KERNEL)) {
    const int cpc = (threads / 32) * (32 / G);
    const size_t chain_b = (size_t)cpc * cfg.ld * sizeof(double);
    const size_t table_b = (size_t)((P.table_doubles + 1) & ~1) * sizeof(double);
    P.table_in_smem = (table_b + chain_b <= 200 * 1024) ? 1 : 0;
    const size_t smem_b = chain_b + (P.table_in_smem ? table_b : 0);
#define DZ_CASE(g, r) if (G == g && R == r) return dreamzs_launch_st2_##g##_##r(P, threads, smem_b, stream);
    DZ_CASE(4, 1) DZ_CASE(8, 1) DZ_CASE(16, 1) DZ_CASE(32, 1) DZ_CASE(32, 2) DZ_CASE(32, 4) DZ_CASE(32, 8)
#undef DZ_CASE
  }
  const int chains_per_cta = (threads / 32) * (32 / G);
  const size_t chain_bytes = (size_t)chains_per_cta * ((size_t)P.nslots * cfg.ld + 3 * DREAMZS_MAX_MULTITRY) * sizeof(double);
  const size_t table_bytes = (size_t)((P.table_doubles + 1) & ~1) * sizeof(double);
  P.table_in_smem = (table_bytes + chain_bytes <= 200 * 1024) ? 1 : 0;
  const size_t smem = chain_bytes + (P.table_in_smem ? table_bytes : 0);
  if (smem > 227 * 1024) return DREAMZS_E_UNSUPPORTED;
#define DZ_CASE(g, r) if (G == g && R == r) return dreamzs_launch_step_##g##_##r(P, threads, smem, stream);
  DZ_CASE(4, 1) DZ_CASE(8, 1) DZ_CASE(16, 1) DZ_CASE(32, 1) DZ_CASE(32, 2) DZ_CASE(32, 4) DZ_CASE(32, 8)
#undef DZ_CASE
  return DREAMZS_E_UNSUPPORTED;
}

static int all_flat_hint(const dreamzs_config *cfg) { return cfg->flags & DREAMZS_FLAG_ALL_FLAT; }

extern "C" int dreamzs_abi_version(void) { return DREAMZS_ABI_VERSION; }

extern "C" int64_t dreamzs_whiten_doubles(int32_t ld) {
  if (ld < 4 || (ld & 3)) return 0;
  return (int64_t)dreamzs::wwin_ntiles_host(ld) * 32;
}

__global__ void rng_normals_kernel(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t call_no, int nblocks, float *out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  float f[4];
  normal4f(philox4x32((uint32_t)b, (call_no << 3) | ST_NORMAL, iter, chain, (uint32_t)seed, (uint32_t)(seed >> 32)), f);
  out[4 * b] = f[0]; out[4 * b + 1] = f[1]; out[4 * b + 2] = f[2]; out[4 * b + 3] = f[3];
}
extern "C" int dreamzs_rng_normals(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t call_no, int32_t nblocks, float *out,
                                   void *stream) {
  if (nblocks < 0 || !out) return DREAMZS_E_BADARG;
  if (nblocks == 0) return DREAMZS_OK;
  rng_normals_kernel<<<(nblocks + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seed, chain, iter, call_no, nblocks, out);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

extern "C" int dreamzs_init_logp(const dreamzs_config *cfg, const dreamzs_state *st, void *stream) {
  int rc = check_cfg(cfg, st);
  if (rc != DREAMZS_OK) return rc;
  if (cfg->nchains_local == 0) return DREAMZS_OK;
  StepParams P{};
  P.cfg = *cfg; P.st = *st; P.init_only = 1; P.all_flat = all_flat_hint(cfg);
  if (wwin_eligible(P)) {
    StepParams G = P;
    G.st.gauss_L = nullptr; G.st.gauss_Y = nullptr;   // evaluate last_prior / last_like with the generic path ...
    rc = dispatch(G, (cudaStream_t)stream);
    if (rc != DREAMZS_OK) return rc;
    rc = dreamzs_launch_whiten(P, (cudaStream_t)stream);   // ... and derive u = L^T x
    if (rc != DREAMZS_OK || !(st->gauss_Y && st->gauss_Q)) return rc;
  }
  if (gwin_eligible(P)) {
    StepParams G = P;
    G.st.gauss_Y = nullptr;   // evaluate last_prior / last_like with the generic path ...
    rc = dispatch(G, (cudaStream_t)stream);
    if (rc != DREAMZS_OK) return rc;
    return dreamzs_launch_gauss_refresh(P, (cudaStream_t)stream);   // ... and derive y = invC x, Q = x.y
  }
  return dispatch(P, (cudaStream_t)stream);
}

static int step_impl(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr, int64_t iter_begin,
                     int32_t niter, int64_t archive_rows, const dreamzs_peers *peers, uint64_t wait_k, uint64_t publish_k,
                     void *stream, const double *temperature = nullptr, bool multi = false, uint64_t k0 = 0);

extern "C" int dreamzs_step(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr,
                            int64_t iter_begin, int32_t niter, int64_t archive_rows, void *stream) {
  return step_impl(cfg, st, tr, iter_begin, niter, archive_rows, nullptr, 0, 0, stream);
}

// astep(q0, T, last_loglike, last_logprior) at per-chain temperatures (Dream.py:193; the tempering driver calls it
// once per chain and iteration, core.py:173, 232): one iteration on the generic kernel, which carries T.
extern "C" int dreamzs_step_tempered(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr,
                                     int64_t iter, int64_t archive_rows, const double *temperature, void *stream) {
  if (!temperature) return DREAMZS_E_BADARG;
  return step_impl(cfg, st, tr, iter, 1, archive_rows, nullptr, 0, 0, stream, temperature);
}

static int64_t appends_between(int64_t t0, int64_t n, int64_t thin) {
  if (n <= 0) return 0;
  const int64_t first = ((t0 + thin - 1) / thin) * thin, last = t0 + n - 1;
  return first > last ? 0 : (last - first) / thin + 1;
}

// multi: the launch spans several windows (persistent whitened window kernel, dreamzs_state.sync_ws); k0 = appends made
// before it
static int step_impl(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr, int64_t iter_begin,
                     int32_t niter, int64_t archive_rows, const dreamzs_peers *peers, uint64_t wait_k, uint64_t publish_k,
                     void *stream, const double *temperature, bool multi, uint64_t k0) {
  int rc = check_cfg(cfg, st);
  if (rc != DREAMZS_OK) return rc;
  if (cfg->target_kind == DREAMZS_TARGET_EXTERNAL) return DREAMZS_E_UNSUPPORTED;   // use dreamzs_propose / dreamzs_accept
  if (!tr || !tr->trace || !tr->trace_logp || niter < 0 || iter_begin < 0) return DREAMZS_E_BADARG;
  if (tr->trace_offset < 0 || tr->trace_offset + niter > tr->trace_iters) return DREAMZS_E_BADARG;
  if (archive_rows < 2 * cfg->nDEpairs || archive_rows > st->Z_capacity_rows) return DREAMZS_E_BADARG;
  if (!multi) {
    // only the last iteration of a launch may append (the archive is read-only inside a launch)
    for (int it = 0; it + 1 < niter; ++it)
      if ((iter_begin + it) % cfg->history_thin == 0) return DREAMZS_E_BADARG;
  } else if (!st->sync_ws || 2 * (appends_between(iter_begin, niter, cfg->history_thin) + 2) + 16 + DREAMZS_SYNC_GROUP_WORDS > st->sync_ws_words) return DREAMZS_E_BADARG;
  if (archive_rows + appends_between(iter_begin, niter, cfg->history_thin) * cfg->nchains_global > st->Z_capacity_rows)
    return DREAMZS_E_BADARG;
  if (niter == 0 || cfg->nchains_local == 0) return DREAMZS_OK;
  StepParams P{};
  P.cfg = *cfg; P.st = *st; P.tr = *tr; P.iter_begin = iter_begin; P.niter = niter; P.archive_rows = archive_rows;
  P.all_flat = all_flat_hint(cfg);
  if (multi) {
    P.ww_sync = st->sync_ws; P.ww_k0 = k0;
    P.ww_wcap = (int32_t)appends_between(iter_begin, niter, cfg->history_thin) + 2;
    // abort word + two counters per appending window of the span (chains appended / forwarded to the peers) + the
    // per-group counters at the end, zeroed in stream order
    const size_t words = (size_t)st->sync_ws_words;
    if (cudaMemsetAsync(st->sync_ws, 0, words * sizeof(uint32_t), (cudaStream_t)stream) != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  }
  if (temperature) {   // only the generic kernel scales the log-likelihood; the dense-Gaussian kernels assume T = 1
    P.temperature = temperature;
    P.cfg.flags |= DREAMZS_FLAG_GENERIC_KERNEL;
    P.st.gauss_Y = nullptr; P.st.gauss_Q = nullptr; P.st.gauss_L = nullptr; P.st.gauss_U = nullptr;
  }
  if (peers) {
    if (peers->world < 1 || peers->world > DREAMZS_MAX_PEERS || peers->rank < 0 || peers->rank >= peers->world) return DREAMZS_E_BADARG;
    for (int q = 0; q < peers->world; ++q)
      if (q != peers->rank) {
        if (!peers->Z[q] || !peers->flags[q]) return DREAMZS_E_BADARG;
        P.peer_flag[P.npeers] = peers->flags[q] + peers->rank;
        P.peer_Z[P.npeers++] = peers->Z[q];
      }
    if (!peers->flags[peers->rank] || !peers->counter || !peers->error) return DREAMZS_E_BADARG;
    P.my_flags = peers->flags[peers->rank]; P.my_rank = peers->rank; P.world = peers->world;
    if (multi && peers->gflag_stride > 0) {
      P.gflag_stride = peers->gflag_stride;
      P.my_pub = peers->flags[peers->rank] + DREAMZS_GFLAG_OFFSET + (size_t)peers->rank * peers->gflag_stride;
      int np = 0;
      for (int q = 0; q < peers->world; ++q)
        if (q != peers->rank) P.peer_pub[np++] = peers->flags[q] + DREAMZS_GFLAG_OFFSET + (size_t)q * peers->gflag_stride;
    }
    P.wait_k = wait_k; P.publish_k = publish_k; P.peer_counter = peers->counter; P.peer_error = peers->error;
  }
  return dispatch(P, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- replicas of the archive over NVLink
extern "C" int dreamzs_shared_alloc(int64_t bytes, void **dev_ptr, void *handle) {
  if (bytes <= 0 || !dev_ptr || !handle) return DREAMZS_E_BADARG;
  void *p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle, p) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(p);
    return DREAMZS_E_LAUNCH;
  }
  *dev_ptr = p;
  return DREAMZS_OK;
}
extern "C" int dreamzs_shared_open(const void *handle, void **dev_ptr) {
  if (!handle || !dev_ptr) return DREAMZS_E_BADARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  if (cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  return DREAMZS_OK;
}
extern "C" int dreamzs_shared_close(void *dev_ptr) {
  if (cudaIpcCloseMemHandle(dev_ptr) != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  return DREAMZS_OK;
}
extern "C" int dreamzs_shared_free(void *dev_ptr) {
  if (cudaFree(dev_ptr) != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  return DREAMZS_OK;
}

// _sample_dream's loop (pydream/core.py:103-122) for the steady state (no adaptation): one launch per window.
extern "C" int dreamzs_run(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr,
                           int64_t iter_begin, int64_t niter, int64_t archive_rows, int64_t appends_done,
                           const dreamzs_peers *peers, dreamzs_append_hook hook, void *user, const dreamzs_adapt *adapt,
                           void *stream, int64_t *launches, int64_t *archive_rows_out) {
  if (!cfg || niter < 0 || iter_begin < 0 || !tr || appends_done < 0) return DREAMZS_E_BADARG;
  if (cfg->history_thin < 1) return DREAMZS_E_BADARG;
  if (peers && (peers->world < 1 || peers->world > DREAMZS_MAX_PEERS || !peers->error)) return DREAMZS_E_BADARG;
  const bool sharded = cfg->nchains_local != cfg->nchains_global;
  if (sharded && !hook && !(peers && peers->world > 1)) return DREAMZS_E_BADARG;   // other shards' rows need a transport
  const bool p2p = peers && peers->world > 1;
  const int64_t thin = cfg->history_thin, end = iter_begin + niter;
  dreamzs_trace w = *tr;
  int64_t t = iter_begin, nl = 0;
  bool waited = appends_done == 0;       // nothing of the peers to wait for before the first append
  const bool adapting = adapt && (adapt->adapt_crossover || adapt->adapt_gamma);
  if (adapting) {
    if (!tr->decisions || !adapt->colsum || !adapt->colsq || !adapt->partial || !adapt->workspace || !adapt->x_entry ||
        !adapt->cr_probs || !adapt->gamma_probs || (sharded && !adapt->reduce))
      return DREAMZS_E_BADARG;
    if (t <= adapt->crossover_burnin &&
        cudaMemcpyAsync(adapt->x_entry, st->X, (size_t)cfg->nchains_local * cfg->ld * sizeof(double), cudaMemcpyDeviceToDevice,
                        (cudaStream_t)stream) != cudaSuccess) {
      (void)cudaGetLastError();
      return DREAMZS_E_LAUNCH;
    }
  }
  // whitened window kernel with scratch words: ONE persistent launch per span of windows (up to the scratch size)
  bool persistent = false;
  if (st->sync_ws && st->sync_ws_words > 32 + DREAMZS_SYNC_GROUP_WORDS && !hook) {
    StepParams Q{};
    Q.cfg = *cfg; Q.st = *st; Q.all_flat = all_flat_hint(cfg);
    persistent = check_cfg(cfg, st) == DREAMZS_OK && wwin_eligible(Q);
  }
  // kernels a window launches: the two-stage multi-try step is three (scalar draws, points, chains)
  int kern_per_window = 1;
  {
    StepParams Q{};
    Q.cfg = *cfg; Q.st = *st;
    if (check_cfg(cfg, st) == DREAMZS_OK && cfg->target_kind != DREAMZS_TARGET_EXTERNAL && mtp_eligible(Q) && st->draw_ws &&
        st->draw_ws_bytes >= dreamzs_draw_ws_bytes(cfg, (int32_t)(thin < niter ? thin : niter)))
      kern_per_window = 3;
  }
  int st2_nb = 0;    // two-stage single-try step: iterations per (draw, chain) kernel pair, 0 = not in use
  {
    StepParams Q{};
    Q.cfg = *cfg; Q.st = *st; Q.all_flat = all_flat_hint(cfg);
    if (check_cfg(cfg, st) == DREAMZS_OK && cfg->multitry == 1 && cfg->target_kind != DREAMZS_TARGET_EXTERNAL && st->draw_ws &&
        st->draw_ws_bytes >= st2_iter_bytes(*cfg) && !(cfg->flags & DREAMZS_FLAG_GENERIC_KERNEL) && !wwin_eligible(Q) &&
        !gwin_eligible(Q) && !(cfg->target_kind == DREAMZS_TARGET_GAUSSIAN_DENSE && cfg->ld > 64 && cfg->ld <= 128))
      st2_nb = (int)(st->draw_ws_bytes / st2_iter_bytes(*cfg));
  }
  while (t < end) {
    const int64_t nxt = ((t + thin - 1) / thin) * thin;          // first appending iteration >= t
    int64_t n = (end < nxt + 1 ? end : nxt + 1) - t;
    const bool burn = adapting && t <= adapt->crossover_burnin;
    if (burn) n = 1;
    if (persistent && !burn) {
      // as many whole windows as the scratch words allow
      const int64_t maxwin = (st->sync_ws_words - DREAMZS_SYNC_GROUP_WORDS - 32) / 2 - 2;
      int64_t span = end - t;
      if (appends_between(t, span, thin) > maxwin) span = (nxt + (maxwin - 1) * thin + 1) - t;
      const int64_t napp = appends_between(t, span, thin);
      w.trace_offset = tr->trace_offset + (t - iter_begin);
      int rc = step_impl(cfg, st, &w, t, (int32_t)span, archive_rows, p2p ? peers : nullptr,
                         (p2p && !waited) ? (uint64_t)appends_done : 0, 0, stream, nullptr, true, (uint64_t)appends_done);
      if (rc != DREAMZS_OK) return rc;
      ++nl;
      waited = napp == 0 ? true : false;      // the next launch waits for the peers' last append of this one
      if (!p2p) waited = true;
      appends_done += napp;
      archive_rows += napp * cfg->nchains_global;
      t += span;
      continue;
    }
    w.trace_offset = tr->trace_offset + (t - iter_begin);
    const bool appends = (t + n - 1) % thin == 0;
    // with peers the launch itself waits for append #appends_done of the others (once) and publishes its own
    int rc = step_impl(cfg, st, &w, t, (int32_t)n, archive_rows, p2p ? peers : nullptr,
                       (p2p && !waited) ? (uint64_t)appends_done : 0, (p2p && appends) ? (uint64_t)(appends_done + 1) : 0, stream);
    if (rc != DREAMZS_OK) return rc;
    waited = true;
    nl += (st2_nb > 0 && !burn) ? 2 * ((n + st2_nb - 1) / st2_nb) : kern_per_window;
    if (appends) {                                               // record_history for every chain (Dream.py:919-938)
      ++appends_done;
      if (p2p) {
        waited = false;
      } else if (hook) {
        rc = hook(user, archive_rows, cfg->nchains_global);
        if (rc != DREAMZS_OK) return rc;
      }
      archive_rows += cfg->nchains_global;
    }
    if (burn && ((10 < t && t < adapt->crossover_burnin) || t == adapt->crossover_burnin)) {
      // one sweep of estimate_crossover_probabilities / estimate_gamma_level_probs (Dream.py:451-540)
      const int64_t trow = w.trace_offset;
      const double *x_old = trow == tr->trace_offset ? adapt->x_entry : tr->trace + (trow - 1) * cfg->ld;
      const int64_t ld_old = trow == tr->trace_offset ? cfg->ld : tr->trace_iters * cfg->ld;
      const int np = 2 * cfg->nCR + 2 * cfg->ngamma;
      rc = dreamzs_adapt_colsum(cfg, st->X, adapt->colsum, adapt->workspace, stream);
      if (rc == DREAMZS_OK && adapt->reduce) rc = adapt->reduce(adapt->user, adapt->colsum, cfg->ndim);
      if (rc == DREAMZS_OK) rc = dreamzs_adapt_colsq(cfg, st->X, adapt->colsum, adapt->colsq, adapt->workspace, stream);
      if (rc == DREAMZS_OK && adapt->reduce) rc = adapt->reduce(adapt->user, adapt->colsq, cfg->ndim);
      if (rc == DREAMZS_OK)
        rc = dreamzs_adapt_jumps(cfg, st->X, x_old, ld_old, tr->decisions + trow, tr->trace_iters, adapt->colsq,
                                 t == adapt->crossover_burnin ? 1 : 0, adapt->adapt_crossover, adapt->adapt_gamma, adapt->partial,
                                 adapt->workspace, stream);
      if (rc == DREAMZS_OK && adapt->reduce) rc = adapt->reduce(adapt->user, adapt->partial, np);
      if (rc == DREAMZS_OK)
        rc = dreamzs_adapt_finish(cfg, adapt->partial, adapt->adapt_crossover, adapt->adapt_gamma, adapt->ncr_updates, adapt->delta_m,
                                  adapt->cr_probs, adapt->ngamma_updates, adapt->delta_m_gamma, adapt->gamma_probs, stream);
      if (rc != DREAMZS_OK) return rc;
      nl += 7;
    }
    t += n;
  }
  if (launches) *launches = nl;
  if (archive_rows_out) *archive_rows_out = archive_rows;
  return DREAMZS_OK;
}

// sampled_params / log_ps leave the device (core.py:81-86): strided device block -> strided host block
extern "C" int dreamzs_copy_d2h_2d(void *dst_host, int64_t dst_pitch_bytes, const void *src_device, int64_t src_pitch_bytes,
                                   int64_t width_bytes, int64_t height, void *stream) {
  if (!dst_host || !src_device || width_bytes < 0 || height < 0 || dst_pitch_bytes < width_bytes || src_pitch_bytes < width_bytes)
    return DREAMZS_E_BADARG;
  if (width_bytes == 0 || height == 0) return DREAMZS_OK;
  const cudaError_t e = cudaMemcpy2DAsync(dst_host, (size_t)dst_pitch_bytes, src_device, (size_t)src_pitch_bytes, (size_t)width_bytes,
                                          (size_t)height, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  return DREAMZS_OK;
}

// ---------------------------------------------------------------- split step for caller-evaluated likelihoods
// Model.total_logp calls the user's likelihood (pydream/model.py:30): with target_kind EXTERNAL one iteration is
// dreamzs_propose -> caller evaluates log L of the proposals on the device -> dreamzs_accept.
static int ext_common(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows, const double *proposals,
                      const double *aux) {
  int rc = check_cfg(cfg, st);
  if (rc != DREAMZS_OK) return rc;
  if (cfg->target_kind != DREAMZS_TARGET_EXTERNAL || !proposals || !aux || iter < 0) return DREAMZS_E_BADARG;
  if (archive_rows < 2 * cfg->nDEpairs || archive_rows > st->Z_capacity_rows) return DREAMZS_E_BADARG;
  return DREAMZS_OK;
}

extern "C" int dreamzs_propose(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows,
                               double *proposals, double *aux, void *stream) {
  int rc = ext_common(cfg, st, iter, archive_rows, proposals, aux);
  if (rc != DREAMZS_OK) return rc;
  if (cfg->nchains_local == 0) return DREAMZS_OK;
  StepParams P{};
  P.cfg = *cfg; P.st = *st; P.iter_begin = iter; P.niter = 1; P.archive_rows = archive_rows;
  P.all_flat = all_flat_hint(cfg);
  P.ext_phase = 1; P.ext_prop = proposals; P.ext_aux = aux;
  return dispatch(P, (cudaStream_t)stream);
}

extern "C" int dreamzs_select(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows,
                              double *proposals, double *aux, const double *loglike, int32_t *error, void *stream) {
  int rc = ext_common(cfg, st, iter, archive_rows, proposals, aux);
  if (rc != DREAMZS_OK) return rc;
  if (cfg->multitry < 2 || !loglike || !error) return DREAMZS_E_BADARG;
  if (cfg->nchains_local == 0) return DREAMZS_OK;
  StepParams P{};
  P.cfg = *cfg; P.st = *st; P.iter_begin = iter; P.niter = 1; P.archive_rows = archive_rows;
  P.all_flat = all_flat_hint(cfg);
  P.ext_phase = 2; P.ext_prop = proposals; P.ext_aux = aux; P.ext_like = loglike; P.ext_error = error;
  return dispatch(P, (cudaStream_t)stream);
}

extern "C" int dreamzs_repropose(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows,
                                 double *proposals, double *aux, const double *loglike, int32_t *count, void *stream) {
  int rc = ext_common(cfg, st, iter, archive_rows, proposals, aux);
  if (rc != DREAMZS_OK) return rc;
  if (cfg->multitry < 2 || !loglike || !count) return DREAMZS_E_BADARG;
  if (cfg->nchains_local == 0) return DREAMZS_OK;
  StepParams P{};
  P.cfg = *cfg; P.st = *st; P.iter_begin = iter; P.niter = 1; P.archive_rows = archive_rows;
  P.all_flat = all_flat_hint(cfg);
  P.ext_phase = 4; P.ext_prop = proposals; P.ext_aux = aux; P.ext_like = loglike; P.ext_error = count;
  return dispatch(P, (cudaStream_t)stream);
}

extern "C" int dreamzs_accept(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr, int64_t iter,
                              int64_t archive_rows, const double *proposals, const double *aux, const double *loglike,
                              void *stream) {
  int rc = ext_common(cfg, st, iter, archive_rows, proposals, aux);
  if (rc != DREAMZS_OK) return rc;
  if (!loglike) return DREAMZS_E_BADARG;
  if (!tr || !tr->trace || !tr->trace_logp || tr->trace_offset < 0 || tr->trace_offset + 1 > tr->trace_iters) return DREAMZS_E_BADARG;
  if (iter % cfg->history_thin == 0 && archive_rows + cfg->nchains_global > st->Z_capacity_rows) return DREAMZS_E_BADARG;
  if (cfg->nchains_local == 0) return DREAMZS_OK;
  StepParams P{};
  P.cfg = *cfg; P.st = *st; P.tr = *tr; P.iter_begin = iter; P.niter = 1; P.archive_rows = archive_rows;
  P.all_flat = all_flat_hint(cfg);
  P.ext_phase = cfg->multitry > 1 ? 3 : 2;
  P.ext_prop = const_cast<double *>(proposals); P.ext_aux = const_cast<double *>(aux); P.ext_like = loglike;
  return dispatch(P, (cudaStream_t)stream);
}
