// Shared device helpers of the MT-DREAM(ZS) kernels: the Philox RNG contract (DESIGN.md
// "RNG contract"; identical to oracle/philox.py), numpy-compatible scalar helpers and the
// lane-group reductions.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/dreamzs.h"

namespace dreamzs {

enum { ST_MULTINOMIAL = 0, ST_SAMPLE, ST_NORMAL, ST_UNIFORM_VEC, ST_UNIFORM_SCAL, ST_RAND, ST_RANDINT };

// decision word layout (include/dreamzs.h dreamzs_trace.decisions)
__host__ __device__ inline uint32_t pack_decision(int changed, int snooker, int cr, int lvl, int delta, int sel,
                                                  int gamma_one, int accepted) {
  return (uint32_t)changed | ((uint32_t)snooker << 1) | ((uint32_t)cr << 2) | ((uint32_t)lvl << 6) |
         ((uint32_t)delta << 10) | ((uint32_t)sel << 14) | ((uint32_t)gamma_one << 18) | ((uint32_t)accepted << 19);
}

// Philox4x32-10 (Salmon et al., SC'11).  Round keys are bumped in registers; the multiplies are
// two IMAD.WIDE per round.  Kept out of line: the step kernels call it from ~20 sites and are bound by
// instruction fetch (same-box A/B: banana d=200 +12-25 %, Gaussian d=50 +6 %, C2 window kernel unchanged).
static __device__ __noinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Per (seed, chain, iteration) stream with running call numbers per primitive.
struct Stream {
  uint32_t k0, k1, chain, iter;
  uint32_t n_multinomial, n_sample, n_normal, n_uvec, n_uscal, n_rand, n_randint;
  __device__ __forceinline__ void init(uint64_t seed, uint32_t chain_, uint32_t iter_) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); chain = chain_; iter = iter_;
    n_multinomial = n_sample = n_normal = n_uvec = n_uscal = n_rand = n_randint = 0;
  }
  __device__ __forceinline__ uint4 block(uint32_t call_no, int st, uint32_t b) const {
    return philox4x32(b, (call_no << 3) | (uint32_t)st, iter, chain, k0, k1);
  }
};

__device__ __forceinline__ double u53_of(uint32_t w0, uint32_t w1) {
  return (double)(((uint64_t)(w0 >> 5) << 26) + (uint64_t)(w1 >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double u32_of(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

// np.random.multinomial(1, p): inverse CDF on a running sum (oracle/philox.py multinomial_index)
__device__ __forceinline__ int multinomial_index(Stream &s, const double *p, int n) {
  const uint4 w = s.block(s.n_multinomial++, ST_MULTINOMIAL, 0);
  const double u = u53_of(w.x, w.y);
  double acc = 0.0;
  int idx = n - 1;
  bool found = false;
  for (int j = 0; j < n; ++j) {
    acc = acc + p[j];
    if (!found && u < acc) { idx = j; found = true; }
  }
  return idx;
}
__device__ __forceinline__ int multinomial2(Stream &s, double p0) {  // [p0, 1-p0]
  const uint4 w = s.block(s.n_multinomial++, ST_MULTINOMIAL, 0);
  const double u = u53_of(w.x, w.y);
  return (u < 0.0 + p0) ? 0 : 1;
}
__device__ __forceinline__ double uniform_scalar(Stream &s) {
  const uint4 w = s.block(s.n_uscal++, ST_UNIFORM_SCAL, 0);
  return u53_of(w.x, w.y);
}

__device__ __forceinline__ double nan_to_num(double x) {
  if (isnan(x)) return 0.0;
  if (isinf(x)) return x > 0 ? DBL_MAX : -DBL_MAX;
  return x;
}

// Normal variates of the RNG contract (stream 2; recipe and rationale in oracle/philox.py normal_pairs32): Box-Muller
// in float32 on 24-bit uniforms, written with operations IEEE-754 rounds identically on every machine (float add,
// multiply, fma, sqrt with explicit round-to-nearest intrinsics, fixed Cephes coefficients), so that the kernels, the
// C oracle and the numpy shim produce the same bits.  zeta (Dream.py:694) is N(0, 1e-12): float32 resolution is ample.
__device__ __forceinline__ void normal_pair32(uint32_t w0, uint32_t w1, float &n0, float &n1) {
  const float a = (float)((w0 >> 8) + 1u);                       // exact: <= 2^24
  const uint32_t bits = __float_as_uint(a);
  int E = (int)(bits >> 23) - 127;
  float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421356f) { m = __fmul_rn(m, 0.5f); E += 1; }
  const float t = __fadd_rn(m, -1.0f), z = __fmul_rn(t, t);
  float P = 7.0376836292E-2f;
  P = __fmaf_rn(P, t, -1.1514610310E-1f); P = __fmaf_rn(P, t, 1.1676998740E-1f); P = __fmaf_rn(P, t, -1.2420140846E-1f);
  P = __fmaf_rn(P, t, 1.4249322787E-1f); P = __fmaf_rn(P, t, -1.6668057665E-1f); P = __fmaf_rn(P, t, 2.0000714765E-1f);
  P = __fmaf_rn(P, t, -2.4999993993E-1f); P = __fmaf_rn(P, t, 3.3333331174E-1f);
  float y = __fmul_rn(t, __fmul_rn(z, P));
  y = __fmaf_rn(-0.5f, z, y);
  const float logm = __fadd_rn(t, y);
  const float L = __fmaf_rn((float)(E - 24), 0.6931471805599453f, logm);
  const float val = __fmul_rn(-2.0f, L);
  const float r = __fsqrt_rn(val > 0.0f ? val : 0.0f);
  const int32_t w24 = (int32_t)(w1 >> 8);
  const int32_t k = (w24 + (1 << 21)) >> 22;
  const float g = __fmul_rn((float)(w24 - (k << 22)), 5.9604644775390625e-08f);   // 2^-24: exact
  const float phi = __fmul_rn(g, 6.283185307179586f);
  const float zz = __fmul_rn(phi, phi);
  const float sp = __fmaf_rn(__fmaf_rn(-1.9515295891E-4f, zz, 8.3321608736E-3f), zz, -1.6666654611E-1f);
  const float cp = __fmaf_rn(__fmaf_rn(2.443315711809948E-5f, zz, -1.388731625493765E-3f), zz, 4.166664568298827E-2f);
  const float s = __fmaf_rn(__fmul_rn(phi, zz), sp, phi);
  const float c = __fmaf_rn(__fmul_rn(zz, zz), cp, __fmaf_rn(-0.5f, zz, 1.0f));
  const int q = k & 3;
  const float cs = q == 0 ? c : q == 1 ? -s : q == 2 ? -c : s;
  const float sn = q == 0 ? s : q == 1 ? c : q == 2 ? -s : -c;
  n0 = __fmul_rn(r, cs); n1 = __fmul_rn(r, sn);
}
// the four words of one block -> four normals (float32 values; widened exactly where a double is wanted)
__device__ __forceinline__ void normal4f(const uint4 w, float out[4]) {
  normal_pair32(w.x, w.y, out[0], out[1]);
  normal_pair32(w.z, w.w, out[2], out[3]);
}
__device__ __forceinline__ void normal4(const uint4 w, double out[4]) {
  float f[4];
  normal4f(w, f);
  out[0] = (double)f[0]; out[1] = (double)f[1]; out[2] = (double)f[2]; out[3] = (double)f[3];
}

// lane-group all-reduce (butterfly; every lane of the group ends with the same bits)
template <int G>
__device__ __forceinline__ double gsum(double v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ int gsum_int(int v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// ---------------------------------------------------------------- replicas of the archive over NVLink (dreamzs_peers)
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// One thread: wait until every peer has published append #k (its rows are then in this rank's replica).
// Returns false (and sets *error) when a peer has not published within DREAMZS_PEER_TIMEOUT_NS.
__device__ __forceinline__ bool peer_wait(const uint64_t *my_flags, int world, int rank, uint64_t k, int32_t *error) {
  const uint64_t t0 = globaltimer_ns();
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    while (ld_acquire_sys(my_flags + q) < k) {
      if (globaltimer_ns() - t0 > DREAMZS_PEER_TIMEOUT_NS) { atomicExch(error, 1); return false; }
      __nanosleep(100);
    }
  }
  return true;
}
// The leader lane of a chain, after the chain's rows went to every replica and every storing lane executed
// __threadfence_system(): count the chain; the last one publishes "append #k done" to every peer.
__device__ __forceinline__ void peer_chain_appended(unsigned int *counter, unsigned int nchains, uint64_t *const *peer_flag,
                                                    int npeers, uint64_t k) {
  const unsigned int old = atomicAdd(counter, 1u);
  if (old == nchains - 1u) {
    atomicExch(counter, 0u);
    __threadfence_system();
    for (int q = 0; q < npeers; ++q) st_release_sys(peer_flag[q], k);
  }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel, device): `cache` is a per-kernel array
// indexed by device ordinal (the launchers keep it in a function-local static).
template <typename K>
inline int ensure_dynamic_smem(K kern, size_t smem, size_t (&cache)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return DREAMZS_E_LAUNCH; }
  dev &= 63;
  if (smem > cache[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      (void)cudaGetLastError();
      return DREAMZS_E_LAUNCH;
    }
    cache[dev] = smem;
  }
  return DREAMZS_OK;
}

}  // namespace dreamzs
