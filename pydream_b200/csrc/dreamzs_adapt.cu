// Crossover-probability / gamma-level adaptation kernels (burn-in only), sm_100a.
// Restates estimate_crossover_probabilities (pydream/Dream.py:451-499) and
// estimate_gamma_level_probs (:501-540) for a synchronous sweep: the column standard deviation
// of the current positions (set_current_position_arr, :424-449) is the same for every chain of
// the sweep, so it is computed once (two-pass, like np.std) and the per-chain squared
// normalised jumps are reduced per CR / gamma-level index in a fixed (deterministic) order.
// Every stage is a plain reduction kernel; partial sums cross GPUs through the caller's
// all-reduce between the stages (include/dreamzs.h).
#include "dreamzs_common.cuh"

namespace dreamzs {

constexpr int ADAPT_ROWS_PER_CTA = 64;
constexpr int ADAPT_TX = 32, ADAPT_TY = 8;

// stage A: partial[cta][i] = sum over the CTA's chains of f(X[c][i]);  f = x  or (x - mean)^2.
// threadIdx.x over dimensions (coalesced rows), threadIdx.y over the CTA's chains; fixed-order tree over y.
__global__ void __launch_bounds__(ADAPT_TX * ADAPT_TY) adapt_col_partial(const double *X, int nchains, int d, int ld, const double *colsum,
                                                                         int nglobal, double *partial) {
  __shared__ double sh[ADAPT_TY][ADAPT_TX];
  const int c0 = blockIdx.x * ADAPT_ROWS_PER_CTA;
  const int c1 = min(nchains, c0 + ADAPT_ROWS_PER_CTA);
  for (int ib = 0; ib < d; ib += ADAPT_TX) {
    const int i = ib + threadIdx.x;
    double acc = 0.0;
    if (i < d) {
      if (colsum) {
        const double mean = colsum[i] / (double)nglobal;
#pragma unroll 4
        for (int c = c0 + threadIdx.y; c < c1; c += ADAPT_TY) { const double t = X[(size_t)c * ld + i] - mean; acc += t * t; }
      } else {
#pragma unroll 4
        for (int c = c0 + threadIdx.y; c < c1; c += ADAPT_TY) acc += X[(size_t)c * ld + i];
      }
    }
    sh[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && i < d) {
      double tot = 0.0;
      for (int y = 0; y < ADAPT_TY; ++y) tot += sh[y][threadIdx.x];
      partial[(size_t)blockIdx.x * d + i] = tot;
    }
    __syncthreads();
  }
}
// stage B: out[i] = sum_cta partial[cta][i]  (threadIdx.y over the CTAs' partials, fixed-order tree)
constexpr int ADAPT_FY = 32;
__global__ void __launch_bounds__(ADAPT_TX * ADAPT_FY) adapt_col_finish(const double *partial, int nctas, int d, double *out) {
  __shared__ double sh[ADAPT_FY][ADAPT_TX];
  const int i = blockIdx.x * ADAPT_TX + threadIdx.x;
  double acc = 0.0;
  if (i < d)
    for (int b = threadIdx.y; b < nctas; b += ADAPT_FY) acc += partial[(size_t)b * d + i];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < d) {
    double tot = 0.0;
    for (int y = 0; y < ADAPT_FY; ++y) tot += sh[y][threadIdx.x];
    out[i] = tot;
  }
}

// one warp per chain: change_c = nan_to_num(sum_i ((x_new - x_old)/sd_i)^2) for both sd conventions
__global__ void adapt_jump_kernel(const double *Xn, const double *Xo, int64_t ld_old, int nchains, int d, int ld,
                                  const double *colsq, int nglobal, const uint32_t *dec, int64_t dec_stride, double *jump_cr,
                                  double *jump_g, uint32_t *dec_packed) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nchains) return;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double sd = sqrt(colsq[i] / (double)nglobal);
    const double dx = Xn[(size_t)warp * ld + i] - Xo[(size_t)warp * ld_old + i];
    const double t0 = dx / (sd == 0 ? 1e-12 : sd);   // Dream.py:479
    const double t1 = dx / sd;                       // Dream.py:527 (no zero guard)
    a += t0 * t0; b += t1 * t1;
  }
  a = gsum<32>(a, 0xffffffffu); b = gsum<32>(b, 0xffffffffu);
  if (lane == 0) {
    jump_cr[warp] = nan_to_num(a); jump_g[warp] = nan_to_num(b);
    dec_packed[warp] = dec[(size_t)warp * dec_stride];   // this iteration's decision words, gathered for the reduction
  }
}

// single CTA, ONE pass over the chains: per-index counts and sums (2 nCR + 2 ngamma outputs) accumulated per
// thread, then reduced in a fixed order (warp butterfly, then the 32 warp results in order): deterministic.
constexpr int ADAPT_MAXOUT = 2 * DREAMZS_MAX_NCR + 2 * DREAMZS_MAX_NGAMMA;
__global__ void __launch_bounds__(256) adapt_reduce_kernel(const double *jump_cr, const double *jump_g, const uint32_t *dec,
                                                            int64_t dec_stride, int nchains, int nCR, int ngamma, int final_update,
                                                            int adapt_cr, int adapt_g, double *partial) {
  __shared__ double sh[32][ADAPT_MAXOUT];
  const int nout = 2 * nCR + 2 * ngamma;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[ADAPT_MAXOUT];
#pragma unroll
  for (int o = 0; o < ADAPT_MAXOUT; ++o) acc[o] = 0.0;
  for (int c = threadIdx.x; c < nchains; c += blockDim.x) {
    const uint32_t w = dec[(size_t)c * dec_stride];
    const int snk = (w >> 1) & 1, gone = (w >> 18) & 1;
    const bool use_cr = adapt_cr && (final_update || !gone);
    const bool use_g = adapt_g && (final_update || (!gone && !snk));
    const int mcr = snk ? nCR - 1 : (int)((w >> 2) & 15), mg = (int)((w >> 6) & 15);
    const double jc = jump_cr[c], jg = jump_g[c];
#pragma unroll
    for (int m = 0; m < DREAMZS_MAX_NCR; ++m) {
      const bool hit = use_cr && m == mcr;
      acc[m] += hit ? 1.0 : 0.0;
      acc[DREAMZS_MAX_NCR + m] += hit ? jc : 0.0;
    }
#pragma unroll
    for (int m = 0; m < DREAMZS_MAX_NGAMMA; ++m) {
      const bool hit = use_g && m == mg;
      acc[2 * DREAMZS_MAX_NCR + m] += hit ? 1.0 : 0.0;
      acc[2 * DREAMZS_MAX_NCR + DREAMZS_MAX_NGAMMA + m] += hit ? jg : 0.0;
    }
  }
#pragma unroll
  for (int o = 0; o < ADAPT_MAXOUT; ++o) {
    acc[o] = gsum<32>(acc[o], 0xffffffffu);
    if (lane == 0) sh[warp][o] = acc[o];
  }
  __syncthreads();
  if (threadIdx.x < nout) {
    // output o of the packed layout [count_cr | sum_cr | count_g | sum_g] -> slot of the padded layout
    const int o = threadIdx.x;
    int slot;
    if (o < nCR) slot = o;
    else if (o < 2 * nCR) slot = DREAMZS_MAX_NCR + (o - nCR);
    else if (o < 2 * nCR + ngamma) slot = 2 * DREAMZS_MAX_NCR + (o - 2 * nCR);
    else slot = 2 * DREAMZS_MAX_NCR + DREAMZS_MAX_NGAMMA + (o - 2 * nCR - ngamma);
    double tot = 0.0;
    const int nw = blockDim.x >> 5;
    for (int wv = 0; wv < nw; ++wv) tot += sh[wv][slot];
    partial[o] = tot;
  }
}

// Dream.py:483-495 / 527-538 after the sweep: fold the (all-reduced) partials in, renormalise.
__global__ void adapt_finish_kernel(const double *partial, int nCR, int ngamma, int nglobal, int adapt_cr, int adapt_g,
                                    double *ncr_updates, double *delta_m, double *cr_probs, double *ngamma_updates,
                                    double *delta_m_gamma, double *gamma_probs) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (adapt_cr) {
    double total = 0.0; bool all = true;
    for (int m = 0; m < nCR; ++m) {
      ncr_updates[m] += partial[m]; delta_m[m] = delta_m[m] + partial[nCR + m];
      total += partial[m]; all = all && (delta_m[m] != 0);
    }
    if (total > 0 && all) {
      double p[DREAMZS_MAX_NCR], sum = 0.0;
      for (int m = 0; m < nCR; ++m) { p[m] = (delta_m[m] / ncr_updates[m]) * nglobal; sum = m == 0 ? p[0] : sum + p[m]; }
      for (int m = 0; m < nCR; ++m) cr_probs[m] = p[m] / sum;
    }
  }
  if (adapt_g) {
    const double *pg = partial + 2 * nCR;
    double total = 0.0; bool all = true;
    for (int m = 0; m < ngamma; ++m) {
      ngamma_updates[m] += pg[m]; delta_m_gamma[m] = delta_m_gamma[m] + pg[ngamma + m];
      total += pg[m]; all = all && (delta_m_gamma[m] != 0);
    }
    if (total > 0 && all) {
      double p[DREAMZS_MAX_NGAMMA], sum = 0.0;
      for (int m = 0; m < ngamma; ++m) { p[m] = (delta_m_gamma[m] / ngamma_updates[m]) * nglobal; sum = m == 0 ? p[0] : sum + p[m]; }
      for (int m = 0; m < ngamma; ++m) gamma_probs[m] = p[m] / sum;
    }
  }
}

}  // namespace dreamzs
using namespace dreamzs;

static int nctas_of(const dreamzs_config *cfg) { return (cfg->nchains_local + ADAPT_ROWS_PER_CTA - 1) / ADAPT_ROWS_PER_CTA; }
static int ok(void) { return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH; }

extern "C" int64_t dreamzs_adapt_workspace_bytes(const dreamzs_config *cfg) {
  if (!cfg) return DREAMZS_E_BADARG;
  const int64_t a = (int64_t)(nctas_of(cfg) > 0 ? nctas_of(cfg) : 1) * cfg->ndim, b = 3 * (int64_t)cfg->nchains_local;
  return (int64_t)sizeof(double) * (a > b ? a : b) + 64;
}

static int col_stage(const dreamzs_config *cfg, const double *X, const double *colsum, double *out, void *ws, void *stream) {
  if (!cfg || !X || !out || !ws) return DREAMZS_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = nctas_of(cfg);
  if (n == 0) { cudaMemsetAsync(out, 0, sizeof(double) * cfg->ndim, s); return ok(); }
  adapt_col_partial<<<n, dim3(ADAPT_TX, ADAPT_TY), 0, s>>>(X, cfg->nchains_local, cfg->ndim, cfg->ld, colsum, cfg->nchains_global,
                                                          (double *)ws);
  adapt_col_finish<<<(cfg->ndim + ADAPT_TX - 1) / ADAPT_TX, dim3(ADAPT_TX, ADAPT_FY), 0, s>>>((const double *)ws, n, cfg->ndim, out);
  return ok();
}

extern "C" int dreamzs_adapt_colsum(const dreamzs_config *cfg, const double *X_new, double *colsum, void *workspace, void *stream) {
  return col_stage(cfg, X_new, nullptr, colsum, workspace, stream);
}
extern "C" int dreamzs_adapt_colsq(const dreamzs_config *cfg, const double *X_new, const double *colsum, double *colsq,
                                   void *workspace, void *stream) {
  if (!colsum) return DREAMZS_E_BADARG;
  return col_stage(cfg, X_new, colsum, colsq, workspace, stream);
}
extern "C" int dreamzs_adapt_jumps(const dreamzs_config *cfg, const double *X_new, const double *x_old, int64_t ld_old,
                                   const uint32_t *decisions, int64_t dec_stride, const double *colsq,
                                   int32_t final_update, int32_t adapt_crossover, int32_t adapt_gamma, double *partial,
                                   void *workspace, void *stream) {
  if (!cfg || !X_new || !x_old || !decisions || !colsq || !partial || !workspace) return DREAMZS_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int N = cfg->nchains_local;
  double *jc = (double *)workspace, *jg = jc + N;
  uint32_t *dp = (uint32_t *)(jg + N);
  if (N > 0) adapt_jump_kernel<<<(N * 32 + 127) / 128, 128, 0, s>>>(X_new, x_old, ld_old, N, cfg->ndim, cfg->ld, colsq,
                                                                      cfg->nchains_global, decisions, dec_stride, jc, jg, dp);
  adapt_reduce_kernel<<<1, 256, 0, s>>>(jc, jg, dp, 1, N, cfg->nCR, cfg->ngamma, final_update,
                                         adapt_crossover, adapt_gamma, partial);
  return ok();
}
extern "C" int dreamzs_adapt_finish(const dreamzs_config *cfg, const double *partial, int32_t adapt_crossover,
                                    int32_t adapt_gamma, double *ncr_updates, double *delta_m, double *cr_probs,
                                    double *ngamma_updates, double *delta_m_gamma, double *gamma_probs, void *stream) {
  if (!cfg || !partial || !ncr_updates || !delta_m || !cr_probs || !ngamma_updates || !delta_m_gamma || !gamma_probs)
    return DREAMZS_E_BADARG;
  adapt_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partial, cfg->nCR, cfg->ngamma, cfg->nchains_global,
                                                          adapt_crossover, adapt_gamma, ncr_updates, delta_m, cr_probs,
                                                          ngamma_updates, delta_m_gamma, gamma_probs);
  return ok();
}
