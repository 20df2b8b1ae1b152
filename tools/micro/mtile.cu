// Microbenchmark of the window kernel's product tile (4 rows per lane x NC columns, K = 100) on B200:
// cycles per j-pair for 1..16 warps per SM and three loop shapes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mtile mtile.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_nv(uint32_t a) { double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }

template <int NC, int MODE>
__global__ void __launch_bounds__(512) k(double *out, long long *cyc, int reps) {
  extern __shared__ __align__(16) double sm[];
  const int ld = 100, d2 = 100, nq = 25;
  double *At = sm, *W = sm + d2 * ld;
  for (int i = threadIdx.x; i < d2 * ld + 16 * NC * ld; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[4][NC];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0;
  uint32_t xaddr[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) xaddr[c] = smem_u32(W + (warp * NC + c) * ld);
  const uint32_t hb = ld * 4, rowb = ld * 8;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (lane < nq) {
      uint32_t ap = smem_u32(At + 2 * lane), joff = 0;
      if (MODE == 0) {
#pragma unroll 1
        for (int j = 0; j < d2; j += 2, ap += 2 * rowb, joff += 16) {
          const double2 a0 = lds(ap), b0 = lds(ap + hb), a1 = lds(ap + rowb), b1 = lds(ap + rowb + hb);
          double2 x[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) x[c] = lds(xaddr[c] + joff);
#pragma unroll
          for (int c = 0; c < NC; ++c) { acc[0][c] = fma(a0.x, x[c].x, acc[0][c]); acc[1][c] = fma(a0.y, x[c].x, acc[1][c]); acc[2][c] = fma(b0.x, x[c].x, acc[2][c]); acc[3][c] = fma(b0.y, x[c].x, acc[3][c]); }
#pragma unroll
          for (int c = 0; c < NC; ++c) { acc[0][c] = fma(a1.x, x[c].y, acc[0][c]); acc[1][c] = fma(a1.y, x[c].y, acc[1][c]); acc[2][c] = fma(b1.x, x[c].y, acc[2][c]); acc[3][c] = fma(b1.y, x[c].y, acc[3][c]); }
        }
      } else if (MODE == 1) {   // non-volatile loads, unroll 2: the compiler may hoist
#pragma unroll 2
        for (int j = 0; j < d2; j += 2, ap += 2 * rowb, joff += 16) {
          const double2 a0 = lds_nv(ap), b0 = lds_nv(ap + hb), a1 = lds_nv(ap + rowb), b1 = lds_nv(ap + rowb + hb);
          double2 x[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) x[c] = lds_nv(xaddr[c] + joff);
#pragma unroll
          for (int c = 0; c < NC; ++c) { acc[0][c] = fma(a0.x, x[c].x, acc[0][c]); acc[1][c] = fma(a0.y, x[c].x, acc[1][c]); acc[2][c] = fma(b0.x, x[c].x, acc[2][c]); acc[3][c] = fma(b0.y, x[c].x, acc[3][c]); }
#pragma unroll
          for (int c = 0; c < NC; ++c) { acc[0][c] = fma(a1.x, x[c].y, acc[0][c]); acc[1][c] = fma(a1.y, x[c].y, acc[1][c]); acc[2][c] = fma(b1.x, x[c].y, acc[2][c]); acc[3][c] = fma(b1.y, x[c].y, acc[3][c]); }
        }
      } else {                  // explicit software pipeline
        double2 a0 = lds(ap), b0 = lds(ap + hb), a1 = lds(ap + rowb), b1 = lds(ap + rowb + hb);
        double2 x[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) x[c] = lds(xaddr[c]);
#pragma unroll 1
        for (int j = 0; j < d2; j += 2) {
          const bool more = j + 2 < d2;
          ap += more ? 2 * rowb : 0u; joff += more ? 16u : 0u;
          const double2 na0 = lds(ap), nb0 = lds(ap + hb), na1 = lds(ap + rowb), nb1 = lds(ap + rowb + hb);
          double2 nx[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) nx[c] = lds(xaddr[c] + joff);
#pragma unroll
          for (int c = 0; c < NC; ++c) { acc[0][c] = fma(a0.x, x[c].x, acc[0][c]); acc[1][c] = fma(a0.y, x[c].x, acc[1][c]); acc[2][c] = fma(b0.x, x[c].x, acc[2][c]); acc[3][c] = fma(b0.y, x[c].x, acc[3][c]); }
#pragma unroll
          for (int c = 0; c < NC; ++c) { acc[0][c] = fma(a1.x, x[c].y, acc[0][c]); acc[1][c] = fma(a1.y, x[c].y, acc[1][c]); acc[2][c] = fma(b1.x, x[c].y, acc[2][c]); acc[3][c] = fma(b1.y, x[c].y, acc[3][c]); }
          a0 = na0; b0 = nb0; a1 = na1; b1 = nb1;
#pragma unroll
          for (int c = 0; c < NC; ++c) x[c] = nx[c];
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) s += acc[r][c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NC, int MODE>
void run(double *out, long long *cyc, const char *name) {
  const int reps = 20;
  const size_t smem = (100 * 100 + 16 * NC * 100) * 8;
  cudaFuncSetAttribute(k<NC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int warps : {1, 4, 8, 16}) {
    k<NC, MODE><<<1, warps * 32, smem>>>(out, cyc, reps);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s NC=%d warps %2d: %7.1f cycles per j-pair (%d DFMA per lane), %5.1f FMA/clk/SM\n", name, NC, warps, (double)h / reps / 50.0, 8 * NC,
           (double)warps * 25 * 8 * NC * 50 * reps / h);
  }
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 8 * 512); cudaMalloc(&cyc, 8);
  run<4, 0>(out, cyc, "volatile loads, in order");
  run<4, 1>(out, cyc, "plain loads, unroll 2");
  run<4, 2>(out, cyc, "software pipelined");
  run<6, 0>(out, cyc, "volatile loads, in order");
  run<6, 1>(out, cyc, "plain loads, unroll 2");
  run<6, 2>(out, cyc, "software pipelined");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
