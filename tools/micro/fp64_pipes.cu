// Microbenchmark: do DFMA (fp64 CUDA-core pipe) and DMMA (mma.sync m8n8k4 f64) share a pipe on B200?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu ; run: ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DFMA, odd warps DMMA
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b) {
  const int warp = threadIdx.x >> 5;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  const bool do_mma = MODE == 1 || (MODE == 2 && (warp & 1));
  if (do_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) dmma(acc[i], acc[i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(double *out, int iters, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out; cudaMalloc(&out, sizeof(double) * 256 * sms * 4);
  const int iters = 20000, blocks = sms * 2;   // 16 warps per SM
  // per warp-iteration: DFMA mode = 16 warp-DFMA = 512 FMA; DMMA mode = 8 DMMA = 8*256 = 2048 FMA
  float t0 = run<0>(out, iters, blocks), t1 = run<1>(out, iters, blocks), t2 = run<2>(out, iters, blocks);
  double warps = blocks * 8.0;
  printf("SMs %d\n", sms);
  printf("DFMA only : %.3f ms  -> %.2f TFLOP/s (%.1f FMA/clk/SM @1.965GHz)\n", t0, warps * iters * 512 * 2 / t0 / 1e9, warps * iters * 512 / (t0 * 1e-3) / sms / 1.965e9);
  printf("DMMA only : %.3f ms  -> %.2f TFLOP/s (%.1f FMA/clk/SM)\n", t1, warps * iters * 2048 * 2 / t1 / 1e9, warps * iters * 2048 / (t1 * 1e-3) / sms / 1.965e9);
  printf("mixed     : %.3f ms  (half the warps each; separate pipes if ~max(%.3f, %.3f), shared if ~%.3f)\n", t2, t0 / 2, t1 / 2, (t0 + t1) / 2);
  return 0;
}
