// Gelman-Rubin R-hat (pydream/convergence.py:3-20) as reduction kernels over the device trace,
// sm_100a.  chain_stats: one CTA per chain, threadIdx.x over dimensions (coalesced row reads),
// threadIdx.y over time slices; two passes (mean, then squared deviations) like np.var.
#include "dreamzs_common.cuh"

namespace dreamzs {

constexpr int GR_TX = 32, GR_TY = 8;

__global__ void __launch_bounds__(GR_TX * GR_TY) gr_chain_stats_kernel(const double *trace, int64_t nsamples,
                                                                       int64_t nburnin, int d, int64_t ld,
                                                                       double *chain_mean, double *chain_var) {
  __shared__ double sh[GR_TY][GR_TX];
  const int64_t c = blockIdx.x;
  const double *base = trace + (size_t)c * nsamples * ld;
  const int64_t n = nsamples - nburnin;
  for (int i0 = 0; i0 < d; i0 += GR_TX) {
    const int i = i0 + threadIdx.x;
    double mean = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      double acc = 0.0;
      if (i < d)
        for (int64_t t = nburnin + threadIdx.y; t < nsamples; t += GR_TY) {
          const double v = base[(size_t)t * ld + i];
          if (pass == 0) acc += v; else { const double r = v - mean; acc = fma(r, r, acc); }
        }
      sh[threadIdx.y][threadIdx.x] = acc;
      __syncthreads();
      double tot = 0.0;
      for (int y = 0; y < GR_TY; ++y) tot += sh[y][threadIdx.x];
      __syncthreads();
      if (pass == 0) mean = tot / (double)n;
      else if (threadIdx.y == 0 && i < d) { chain_mean[(size_t)c * d + i] = mean; chain_var[(size_t)c * d + i] = tot / (double)n; }
    }
  }
}

__global__ void gr_finish_kernel(const double *chain_mean, const double *chain_var, int64_t nchains, int64_t nsamples,
                                 int d, double *rhat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  double W = 0.0, mm = 0.0, B = 0.0;
  for (int64_t c = 0; c < nchains; ++c) { W += chain_var[(size_t)c * d + i]; mm += chain_mean[(size_t)c * d + i]; }
  W /= (double)nchains; mm /= (double)nchains;
  for (int64_t c = 0; c < nchains; ++c) { const double r = chain_mean[(size_t)c * d + i] - mm; B = fma(r, r, B); }
  B /= (double)nchains;
  rhat[i] = sqrt((W * (1 - (1. / (double)nsamples)) + B) / W);
}

}  // namespace dreamzs
using namespace dreamzs;

extern "C" int dreamzs_gr_chain_stats(const double *trace, int64_t nchains, int64_t nsamples, int64_t nburnin,
                                      int32_t ndim, int64_t ld, double *chain_mean, double *chain_var, void *stream) {
  if (!trace || !chain_mean || !chain_var || nchains < 0 || nsamples < 1 || nburnin < 0 || nburnin >= nsamples || ndim < 1 || ld < ndim)
    return DREAMZS_E_BADARG;
  if (nchains == 0) return DREAMZS_OK;
  gr_chain_stats_kernel<<<(unsigned)nchains, dim3(GR_TX, GR_TY), 0, (cudaStream_t)stream>>>(trace, nsamples, nburnin, ndim, ld,
                                                                                              chain_mean, chain_var);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

extern "C" int dreamzs_gr_finish(const double *chain_mean, const double *chain_var, int64_t nchains, int64_t nsamples,
                                 int32_t ndim, double *rhat, void *stream) {
  if (!chain_mean || !chain_var || !rhat || nchains < 1 || nsamples < 1 || ndim < 1) return DREAMZS_E_BADARG;
  gr_finish_kernel<<<(ndim + 127) / 128, 128, 0, (cudaStream_t)stream>>>(chain_mean, chain_var, nchains, nsamples, ndim, rhat);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}
