// Gelman-Rubin R-hat (pydream/convergence.py:3-20) as reduction kernels over the device trace, sm_100a.
//
// chain_stats: one CTA per chain streams the chain's post-burn-in block [n, ld] ONCE (HBM-bound:
// 8 d n bytes per chain), 16-B loads, 4 rows in flight per thread.  Mean and variance (ddof 0, as np.var)
// come from sums of the data shifted by the chain's first sample (one pass; the shift keeps the
// cancellation in E[(x-s)^2] - (E[x-s])^2 to a few ulp).  finish: W, B and Rhat per dimension with the
// chains spread over threadIdx.y (fixed-order tree: deterministic).
#include "dreamzs_common.cuh"

namespace dreamzs {

constexpr int GR_THREADS = 256;

__global__ void __launch_bounds__(GR_THREADS) gr_chain_stats_kernel(const double *trace, int64_t nsamples,
                                                                    int64_t nburnin, int d, int64_t ld,
                                                                    double *chain_mean, double *chain_var) {
  extern __shared__ double sh[];            // [rows_per_pass][2 * pairs_per_tile] x 2 (sum, sum of squares)
  const int64_t c = blockIdx.x;
  const double *base = trace + ((size_t)c * nsamples + nburnin) * ld;
  const int64_t n = nsamples - nburnin;
  const int hp = (int)(ld / 2);                              // column pairs per row (ld is a multiple of 4)
  const int hpt = hp < GR_THREADS ? hp : GR_THREADS;         // column pairs per tile of the CTA
  const int rpp = GR_THREADS / hpt;                          // rows handled per pass of the CTA
  const int lp = threadIdx.x % hpt, r0 = threadIdx.x / hpt;
  for (int cpb = 0; cpb < hp; cpb += hpt) {
    const int cp = cpb + lp;
    const bool act = r0 < rpp && cp < hp;
    if (act) {
      double s1x = 0.0, s1y = 0.0, s2x = 0.0, s2y = 0.0;
      const double2 shift = *reinterpret_cast<const double2 *>(base + 2 * cp);
      int64_t t = r0;
      for (; t + 3 * rpp < n; t += 4 * rpp) {
        const double2 v0 = *reinterpret_cast<const double2 *>(base + (size_t)t * ld + 2 * cp);
        const double2 v1 = *reinterpret_cast<const double2 *>(base + (size_t)(t + rpp) * ld + 2 * cp);
        const double2 v2 = *reinterpret_cast<const double2 *>(base + (size_t)(t + 2 * rpp) * ld + 2 * cp);
        const double2 v3 = *reinterpret_cast<const double2 *>(base + (size_t)(t + 3 * rpp) * ld + 2 * cp);
        const double a0 = v0.x - shift.x, a1 = v1.x - shift.x, a2 = v2.x - shift.x, a3 = v3.x - shift.x;
        const double b0 = v0.y - shift.y, b1 = v1.y - shift.y, b2 = v2.y - shift.y, b3 = v3.y - shift.y;
        s1x += (a0 + a1) + (a2 + a3); s1y += (b0 + b1) + (b2 + b3);
        s2x = fma(a0, a0, s2x); s2x = fma(a1, a1, s2x); s2x = fma(a2, a2, s2x); s2x = fma(a3, a3, s2x);
        s2y = fma(b0, b0, s2y); s2y = fma(b1, b1, s2y); s2y = fma(b2, b2, s2y); s2y = fma(b3, b3, s2y);
      }
      for (; t < n; t += rpp) {
        const double2 v = *reinterpret_cast<const double2 *>(base + (size_t)t * ld + 2 * cp);
        const double a = v.x - shift.x, b = v.y - shift.y;
        s1x += a; s1y += b; s2x = fma(a, a, s2x); s2y = fma(b, b, s2y);
      }
      double *p1 = sh + ((size_t)r0 * hpt + lp) * 2, *p2 = sh + (size_t)rpp * hpt * 2 + ((size_t)r0 * hpt + lp) * 2;
      p1[0] = s1x; p1[1] = s1y; p2[0] = s2x; p2[1] = s2y;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * hpt; k += GR_THREADS) {
      const int i = 2 * cpb + k;
      if (i < d) {
        double t1 = 0.0, t2 = 0.0;
        for (int r = 0; r < rpp; ++r) { t1 += sh[(size_t)r * hpt * 2 + k]; t2 += sh[(size_t)rpp * hpt * 2 + (size_t)r * hpt * 2 + k]; }
        const double m1 = t1 / (double)n, m2 = t2 / (double)n;
        chain_mean[(size_t)c * d + i] = base[i] + m1;
        chain_var[(size_t)c * d + i] = m2 - m1 * m1;
      }
    }
    __syncthreads();
  }
}

constexpr int GRF_TX = 32, GRF_TY = 32;

__global__ void __launch_bounds__(GRF_TX * GRF_TY) gr_finish_kernel(const double *chain_mean, const double *chain_var,
                                                                    int64_t nchains, int64_t nsamples, int d, double *rhat) {
  __shared__ double sw[GRF_TY][GRF_TX], s1[GRF_TY][GRF_TX], s2[GRF_TY][GRF_TX];
  const int i = blockIdx.x * GRF_TX + threadIdx.x;
  double W = 0.0, a1 = 0.0, a2 = 0.0;
  const double m0 = i < d ? chain_mean[i] : 0.0;        // shift: the first chain's mean
  if (i < d)
    for (int64_t c = threadIdx.y; c < nchains; c += GRF_TY) {
      W += chain_var[(size_t)c * d + i];
      const double r = chain_mean[(size_t)c * d + i] - m0;
      a1 += r; a2 = fma(r, r, a2);
    }
  sw[threadIdx.y][threadIdx.x] = W; s1[threadIdx.y][threadIdx.x] = a1; s2[threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && i < d) {
    W = 0.0; a1 = 0.0; a2 = 0.0;
    for (int y = 0; y < GRF_TY; ++y) { W += sw[y][threadIdx.x]; a1 += s1[y][threadIdx.x]; a2 += s2[y][threadIdx.x]; }
    W /= (double)nchains;
    const double e1 = a1 / (double)nchains;
    const double B = a2 / (double)nchains - e1 * e1;      // np.var of the chain means, ddof 0
    rhat[i] = sqrt((W * (1 - (1. / (double)nsamples)) + B) / W);
  }
}

}  // namespace dreamzs
using namespace dreamzs;

extern "C" int dreamzs_gr_chain_stats(const double *trace, int64_t nchains, int64_t nsamples, int64_t nburnin,
                                      int32_t ndim, int64_t ld, double *chain_mean, double *chain_var, void *stream) {
  if (!trace || !chain_mean || !chain_var || nchains < 0 || nsamples < 1 || nburnin < 0 || nburnin >= nsamples || ndim < 1 || ld < ndim)
    return DREAMZS_E_BADARG;
  if (nchains == 0) return DREAMZS_OK;
  if (ld & 3) return DREAMZS_E_BADARG;
  const int hp = (int)(ld / 2), hpt = hp < GR_THREADS ? hp : GR_THREADS, rpp = GR_THREADS / hpt;
  const size_t smem = (size_t)2 * rpp * hpt * 2 * sizeof(double);
  gr_chain_stats_kernel<<<(unsigned)nchains, GR_THREADS, smem, (cudaStream_t)stream>>>(trace, nsamples, nburnin, ndim, ld,
                                                                                        chain_mean, chain_var);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

extern "C" int dreamzs_gr_finish(const double *chain_mean, const double *chain_var, int64_t nchains, int64_t nsamples,
                                 int32_t ndim, double *rhat, void *stream) {
  if (!chain_mean || !chain_var || !rhat || nchains < 1 || nsamples < 1 || ndim < 1) return DREAMZS_E_BADARG;
  gr_finish_kernel<<<(ndim + GRF_TX - 1) / GRF_TX, dim3(GRF_TX, GRF_TY), 0, (cudaStream_t)stream>>>(chain_mean, chain_var, nchains,
                                                                                                   nsamples, ndim, rhat);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}
