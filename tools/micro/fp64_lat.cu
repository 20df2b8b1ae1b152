// Microbenchmarks of the B200 fp64 pipe: dependent-op latency and throughput vs. resident warps / ILP.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, long long *cyc, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// all three operands distinct registers per DFMA (no operand reuse): acc[i] = fma(x[i], y[i], acc[i])
template <int ILP>
__global__ void k_dfma3(double *out, long long *cyc, int iters, const double *in) {
  double acc[ILP], x[ILP], y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x * 1e-3 + i; x[i] = in[i]; y[i] = in[ILP + i]; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(x[i], y[(i + it) & (ILP - 1)], acc[i]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_shfl(double *out, long long *cyc, int iters) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_lds(double *out, long long *cyc, int iters) {
  __shared__ int idx[256];
  idx[threadIdx.x] = (threadIdx.x + 1) & 255;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) p = idx[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename F>
void run(const char *name, F f, int threads, int iters, int ops_per_iter, int sms, long long *cyc) {
  f();
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_op = (double)h / iters / ops_per_iter;
  printf("%-34s threads/CTA %4d: %8.2f cycles per warp-op-slot, %6.2f FMA/clk/SM\n", name, threads, per_op,
         (double)threads * ops_per_iter * iters / h);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out, *in; long long *cyc;
  cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&cyc, 64); cudaMalloc(&in, 8 * 256); cudaMemset(in, 0, 8 * 256);
  const int it = 4000;
  for (int th : {32, 128, 256, 512, 1024}) {
    run("dep DFMA ILP1", [&] { k_dfma<1><<<1, th>>>(out, cyc, it, 1.0000001, 1e-9); }, th, it, 1, sms, cyc);
    run("DFMA ILP2", [&] { k_dfma<2><<<1, th>>>(out, cyc, it, 1.0000001, 1e-9); }, th, it, 2, sms, cyc);
    run("DFMA ILP4", [&] { k_dfma<4><<<1, th>>>(out, cyc, it, 1.0000001, 1e-9); }, th, it, 4, sms, cyc);
    run("DFMA ILP8", [&] { k_dfma<8><<<1, th>>>(out, cyc, it, 1.0000001, 1e-9); }, th, it, 8, sms, cyc);
    run("DFMA ILP16", [&] { k_dfma<16><<<1, th>>>(out, cyc, it, 1.0000001, 1e-9); }, th, it, 16, sms, cyc);
    run("DFMA ILP32", [&] { k_dfma<32><<<1, th>>>(out, cyc, it, 1.0000001, 1e-9); }, th, it, 32, sms, cyc);
    run("DFMA 3 distinct operands ILP16", [&] { k_dfma3<16><<<1, th>>>(out, cyc, it, in); }, th, it, 16, sms, cyc);
    run("DFMA 3 distinct operands ILP32", [&] { k_dfma3<32><<<1, th>>>(out, cyc, it, in); }, th, it, 32, sms, cyc);
  }
  run("warp allreduce double (5 shfl+add)", [&] { k_shfl<<<1, 32>>>(out, cyc, it); }, 32, it, 1, sms, cyc);
  run("dependent LDS (pointer chase)", [&] { k_lds<<<1, 32>>>(out, cyc, it); }, 32, it, 1, sms, cyc);
  return 0;
}
