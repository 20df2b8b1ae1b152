// Single-warp dependent-chain latencies on B200 (cycles per op), incl. the building blocks of the sampler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pydream_b200/csrc/dreamzs_common.cuh"
using namespace dreamzs;

#define BENCH(name, init, body, finish)                                             \
  __global__ void name(long long *cyc, double *out, int iters, uint32_t seed) {     \
    init;                                                                           \
    long long t0 = clock64();                                                       \
    for (int it = 0; it < iters; ++it) { body; }                                    \
    long long t1 = clock64();                                                       \
    finish;                                                                         \
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;                      \
  }

BENCH(k_iadd, uint32_t a = seed + threadIdx.x, { a = a * 3u + 7u; }, out[threadIdx.x] = a)
BENCH(k_lop, uint32_t a = seed + threadIdx.x, { a = (a ^ 0x9e3779b9u) + (a >> 3); }, out[threadIdx.x] = a)
BENCH(k_imadw, uint64_t a = seed + threadIdx.x, { a = (uint64_t)(uint32_t)a * 0xD2511F53u + (a >> 32); }, out[threadIdx.x] = (double)a)
BENCH(k_philox, uint4 w = make_uint4(seed, threadIdx.x, 1, 2), { w = philox4x32(w.x, w.y, w.z, w.w, 1, 2); }, out[threadIdx.x] = w.x)
BENCH(k_log, double x = 0.3 + 1e-3 * threadIdx.x, { x = log(x + 1.5) ; }, out[threadIdx.x] = x)
BENCH(k_sqrt, double x = 0.3 + 1e-3 * threadIdx.x, { x = sqrt(x + 1.5); }, out[threadIdx.x] = x)
BENCH(k_scp, double x = 0.3 + 1e-3 * threadIdx.x, { double sc[2]; sincospi(x, &sc[0], &sc[1]); x = sc[0] + sc[1]; }, out[threadIdx.x] = x)
BENCH(k_normal4, uint4 w = make_uint4(seed, threadIdx.x, 1, 2), { double o[4]; normal4(w, o); w.x += (uint32_t)(o[0] + o[1] + o[2] + o[3]); }, out[threadIdx.x] = w.x)
BENCH(k_dadd, double x = 0.3 + 1e-3 * threadIdx.x, { x = x + 1.5; }, out[threadIdx.x] = x)
BENCH(k_dsetp, double x = 0.3 + 1e-3 * threadIdx.x + 1000; int c = 0, { if (x > 0.5 * c) c++; }, out[threadIdx.x] = c)
BENCH(k_u53, uint32_t a = seed + threadIdx.x; double x = 0, { x = u53_of(a, a ^ 5u); a += (uint32_t)(x * 16.0); }, out[threadIdx.x] = x)
BENCH(k_ballot, uint32_t a = seed + threadIdx.x, { a += __ballot_sync(0xffffffffu, a & 1u); }, out[threadIdx.x] = a)
BENCH(k_shfl32, uint32_t a = seed + threadIdx.x, { a += __shfl_xor_sync(0xffffffffu, a, 1); }, out[threadIdx.x] = a)
BENCH(k_redux, int a = seed + threadIdx.x, { a = __reduce_add_sync(0xffffffffu, a) & 0xff; }, out[threadIdx.x] = a)

int main() {
  long long *cyc; double *out;
  cudaMalloc(&cyc, 64); cudaMalloc(&out, 8 * 4096);
  const int it = 2000;
#define RUN(k, threads) { k<<<1, threads>>>(cyc, out, it, 12345u); cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-12s threads %4d: %8.1f cycles/iter\n", #k, threads, (double)h / it); }
  for (int th : {32, 512}) {
    RUN(k_iadd, th) RUN(k_lop, th) RUN(k_imadw, th) RUN(k_philox, th) RUN(k_log, th) RUN(k_sqrt, th) RUN(k_scp, th) RUN(k_normal4, th)
    RUN(k_dadd, th) RUN(k_dsetp, th) RUN(k_u53, th) RUN(k_ballot, th) RUN(k_shfl32, th) RUN(k_redux, th)
  }
  return 0;
}
