// One translation unit per chains-per-CTA variant of the dense-Gaussian window kernel: -DDZ_TC=<chains per CTA>.
#include "dreamzs_gwin_kernel.cuh"
#define DZ_CAT2(a) dreamzs_launch_gwin_##a
#define DZ_CAT(a) DZ_CAT2(a)
int DZ_CAT(DZ_TC)(dreamzs::StepParams &P, cudaStream_t stream) { return dreamzs::launch_gwin<DZ_TC>(P, stream); }

#if DZ_TC == 8
int dreamzs_gwin_usable(const dreamzs_config &cfg, int TC) {
  size_t smem = 0;
  return dreamzs::gwin_pick_nb(cfg, TC, &smem) > 0;
}

// gauss_Y = invC x, gauss_Q = x . invC x for every local chain (one warp per chain, table read through L2);
// run by dreamzs_init_logp so that a window kernel launch may start at any iteration.
__global__ void __launch_bounds__(128) dreamzs_gauss_refresh_kernel(const dreamzs::StepParams P) {
  __shared__ double xs[4][DREAMZS_MAX_NDIM > 128 ? 128 : DREAMZS_MAX_NDIM];
  const int d = P.cfg.ndim, ld = P.cfg.ld, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 4 + warp;
  if (c >= P.cfg.nchains_local) return;
  const int i0 = 4 * lane;
  const bool own = i0 < ld;
  double x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
  if (own) {
    const double *xr = P.st.X + (size_t)c * ld + i0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { x[j] = xr[j]; xs[warp][i0 + j] = x[j]; }
  }
  __syncwarp();
  const double *At = P.st.target_table + 2;
  if (own)
    for (int j = 0; j < d; ++j) {
      const double xj = xs[warp][j];
      const double *a = At + (size_t)j * ld + i0;
#pragma unroll
      for (int r = 0; r < 4; ++r) y[r] = fma(a[r], xj, y[r]);
    }
  double part = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) part = fma(x[j], y[j], part);
  part = dreamzs::gsum<32>(part, 0xffffffffu);
  if (own) {
    double *yr = P.st.gauss_Y + (size_t)c * ld + i0;
#pragma unroll
    for (int j = 0; j < 4; ++j) yr[j] = y[j];
  }
  if (lane == 0) P.st.gauss_Q[c] = part;
}

int dreamzs_launch_gauss_refresh(const dreamzs::StepParams &P, cudaStream_t stream) {
  dreamzs_gauss_refresh_kernel<<<(P.cfg.nchains_local + 3) / 4, 128, 0, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}
#endif
