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

constexpr int ADAPT_ROWS_PER_CTA = 256;

// stage A: partial[cta][i] = sum over the CTA's chains of f(X[c][i]);  f = x  or (x - mean)^2
__global__ void adapt_col_partial(const double *X, int nchains, int d, int ld, const double *colsum, double inv_n_is_div,
                                  int nglobal, double *partial) {
  const int c0 = blockIdx.x * ADAPT_ROWS_PER_CTA;
  const int c1 = min(nchains, c0 + ADAPT_ROWS_PER_CTA);
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    double acc = 0.0;
    if (colsum) {
      const double mean = colsum[i] / (double)nglobal;
      for (int c = c0; c < c1; ++c) { const double t = X[(size_t)c * ld + i] - mean; acc += t * t; }
    } else {
      for (int c = c0; c < c1; ++c) acc += X[(size_t)c * ld + i];
    }
    partial[(size_t)blockIdx.x * d + i] = acc;
  }
}
// stage B: out[i] = sum_cta partial[cta][i]  (fixed order)
__global__ void adapt_col_finish(const double *partial, int nctas, int d, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  double acc = 0.0;
  for (int b = 0; b < nctas; ++b) acc += partial[(size_t)b * d + i];
  out[i] = acc;
}

// one warp per chain: change_c = nan_to_num(sum_i ((x_new - x_old)/sd_i)^2) for both sd conventions
__global__ void adapt_jump_kernel(const double *Xn, const double *Xo, int64_t ld_old, int nchains, int d, int ld,
                                  const double *colsq, int nglobal, double *jump_cr, double *jump_g) {
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
  if (lane == 0) { jump_cr[warp] = nan_to_num(a); jump_g[warp] = nan_to_num(b); }
}

// single CTA: per-index counts and sums over the local chains in a fixed tree order
__global__ void adapt_reduce_kernel(const double *jump_cr, const double *jump_g, const uint32_t *dec, int64_t dec_stride,
                                    int nchains, int nCR, int ngamma, int final_update, int adapt_cr, int adapt_g,
                                    double *partial) {
  __shared__ double sh[1024];
  const int nout = 2 * nCR + 2 * ngamma;
  for (int o = 0; o < nout; ++o) {
    const bool is_cr = o < 2 * nCR;
    const int idx = is_cr ? o % nCR : (o - 2 * nCR) % ngamma;
    const bool is_count = is_cr ? o < nCR : (o - 2 * nCR) < ngamma;
    double acc = 0.0;
    if ((is_cr && adapt_cr) || (!is_cr && adapt_g)) {
      for (int c = threadIdx.x; c < nchains; c += blockDim.x) {
        const uint32_t w = dec[(size_t)c * dec_stride];
        const int snk = (w >> 1) & 1, gone = (w >> 18) & 1;
        bool use; int m;
        if (is_cr) { use = final_update || !gone; m = snk ? nCR - 1 : (int)((w >> 2) & 15); }
        else { use = final_update || (!gone && !snk); m = (int)((w >> 6) & 15); }
        if (use && m == idx) acc += is_count ? 1.0 : (is_cr ? jump_cr[c] : jump_g[c]);
      }
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[o] = sh[0];
    __syncthreads();
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
  const int64_t a = (int64_t)(nctas_of(cfg) > 0 ? nctas_of(cfg) : 1) * cfg->ndim, b = 2 * (int64_t)cfg->nchains_local;
  return (int64_t)sizeof(double) * (a > b ? a : b) + 64;
}

static int col_stage(const dreamzs_config *cfg, const double *X, const double *colsum, double *out, void *ws, void *stream) {
  if (!cfg || !X || !out || !ws) return DREAMZS_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int n = nctas_of(cfg);
  if (n == 0) { cudaMemsetAsync(out, 0, sizeof(double) * cfg->ndim, s); return ok(); }
  adapt_col_partial<<<n, 128, 0, s>>>(X, cfg->nchains_local, cfg->ndim, cfg->ld, colsum, 0.0, cfg->nchains_global, (double *)ws);
  adapt_col_finish<<<(cfg->ndim + 127) / 128, 128, 0, s>>>((const double *)ws, n, cfg->ndim, out);
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
  if (N > 0) adapt_jump_kernel<<<(N * 32 + 127) / 128, 128, 0, s>>>(X_new, x_old, ld_old, N, cfg->ndim, cfg->ld, colsq,
                                                                      cfg->nchains_global, jc, jg);
  adapt_reduce_kernel<<<1, 1024, 0, s>>>(jc, jg, decisions, dec_stride, N, cfg->nCR, cfg->ngamma, final_update,
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
