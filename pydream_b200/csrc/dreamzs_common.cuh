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

// Box-Muller on the four words of one block -> four normals (oracle/philox.py normal_vec)
__device__ __forceinline__ void normal4(const uint4 w, double out[4]) {
#if defined(DZ_FAST_NORMAL) && DZ_FAST_NORMAL
  // A/B builds only (default off; NOT the RNG contract): Box-Muller in float32 (24-bit normals, ~60 instead of ~250
  // issue cycles per block).  With zeta = 1e-12 the difference to the fp64 normals is ~1e-19 on a state of O(1),
  // i.e. below its last bit except for rare rounding flips; to be decided on measurements.
  const float f0 = sqrtf(-2.0f * logf(((float)w.x + 1.0f) * (1.0f / 4294967296.0f)));
  const float f1 = sqrtf(-2.0f * logf(((float)w.z + 1.0f) * (1.0f / 4294967296.0f)));
  float fs0, fc0, fs1, fc1;
  sincospif((float)w.y * (1.0f / 2147483648.0f), &fs0, &fc0);
  sincospif((float)w.w * (1.0f / 2147483648.0f), &fs1, &fc1);
  out[0] = (double)(f0 * fc0); out[1] = (double)(f0 * fs0); out[2] = (double)(f1 * fc1); out[3] = (double)(f1 * fs1);
  return;
#endif
  const double r0 = sqrt(-2.0 * log(((double)w.x + 1.0) * (1.0 / 4294967296.0)));
  const double r1 = sqrt(-2.0 * log(((double)w.z + 1.0) * (1.0 / 4294967296.0)));
  double s0, c0, s1, c1;
  sincospi((double)w.y * (1.0 / 2147483648.0), &s0, &c0);   // 2*pi*u == pi*(2u)
  sincospi((double)w.w * (1.0 / 2147483648.0), &s1, &c1);
  out[0] = r0 * c0; out[1] = r0 * s0; out[2] = r1 * c1; out[3] = r1 * s1;
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
__device__ __forceinline__ void peer_wait(const uint64_t *my_flags, int world, int rank, uint64_t k, int32_t *error) {
  const uint64_t t0 = globaltimer_ns();
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    while (ld_acquire_sys(my_flags + q) < k) {
      if (globaltimer_ns() - t0 > DREAMZS_PEER_TIMEOUT_NS) { atomicExch(error, 1); return; }
      __nanosleep(100);
    }
  }
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
