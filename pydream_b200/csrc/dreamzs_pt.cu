// Temperature swap of the parallel-tempering driver (_sample_dream_pt, pydream/core.py:183-225), sm_100a.
// One iteration of the driver is: astep for every chain at its temperature (dreamzs_step_tempered, recorded in
// trace row 2t), then ONE proposed exchange between two chains drawn by the parent process, then every chain is
// recorded again (trace row 2t+1).  Two stream-ordered launches:
//   decide : one warp serves the driver's draws from the stream of the pseudo-chain 0xFFFFFFFF
//            (np.random.choice(nchains, 2, replace=False) = random.sample contract, np.random.uniform()),
//            evaluates alpha (core.py:193) and parks the pair, the verdict and the two chains' scalars in `ws`
//   apply  : a lane-group per chain copies the state recorded in row 2t (its own, or its partner's when the swap
//            was accepted) into row 2t+1 and, for the two exchanged chains, into X / last_like / last_prior
#include "dreamzs_common.cuh"

namespace dreamzs {

__global__ void pt_decide_kernel(dreamzs_config cfg, const double *last_like, const double *last_prior,
                                 const double *temperature, int64_t iter, double *ws) {
  if (threadIdx.x != 0) return;
  const uint32_t k0 = (uint32_t)cfg.seed, k1 = (uint32_t)(cfg.seed >> 32);
  const uint32_t N = (uint32_t)cfg.nchains_global;
  const uint4 w = philox4x32(0u, (0u << 3) | ST_SAMPLE, (uint32_t)iter, 0xFFFFFFFFu, k0, k1);
  int64_t a = (int64_t)(((uint64_t)w.x * (uint64_t)N) >> 32);
  int64_t b = (int64_t)(((uint64_t)w.y * (uint64_t)(N - 1)) >> 32);
  if (b >= a) b += 1;
  const uint4 wu = philox4x32(0u, (0u << 3) | ST_UNIFORM_SCAL, (uint32_t)iter, 0xFFFFFFFFu, k0, k1);
  const double u = u53_of(wu.x, wu.y);
  const double T1 = temperature[a], T2 = temperature[b], l1 = last_like[a], l2 = last_like[b];
  const double alpha = ((T1 * l2) + (T2 * l1)) - ((T1 * l1) + (T2 * l2));      // core.py:193
  const bool swap = log(u) < alpha;                                          // core.py:195 (false for a nan alpha)
  ws[0] = (double)a; ws[1] = (double)b; ws[2] = swap ? 1.0 : 0.0; ws[3] = alpha;
  ws[4] = l1; ws[5] = last_prior[a]; ws[6] = l2; ws[7] = last_prior[b];
}

// 8 lanes per chain, 16-B accesses
__global__ void __launch_bounds__(256) pt_apply_kernel(dreamzs_config cfg, double *X, double *last_like, double *last_prior,
                                                       dreamzs_trace tr, const double *ws) {
  const int c = blockIdx.x * 32 + (threadIdx.x >> 3), g = threadIdx.x & 7;
  if (c >= cfg.nchains_local) return;
  const int a = (int)ws[0], b = (int)ws[1];
  const bool swap = ws[2] != 0.0;
  const int src = (swap && c == a) ? b : (swap && c == b) ? a : c;
  const int64_t row = tr.trace_offset;          // this launch writes row `row`; row - 1 holds the states after astep
  const int ld = cfg.ld;
  const double2 *from = reinterpret_cast<const double2 *>(tr.trace + ((size_t)src * tr.trace_iters + (row - 1)) * ld);
  double2 *to = reinterpret_cast<double2 *>(tr.trace + ((size_t)c * tr.trace_iters + row) * ld);
  double2 *xr = reinterpret_cast<double2 *>(X + (size_t)c * ld);
  for (int i = g; i < ld / 2; i += 8) {
    const double2 v = from[i];
    to[i] = v;
    if (src != c) xr[i] = v;
  }
  if (g == 0) {
    // logpnews travels with the state (core.py:207-208): T of the chain that produced it
    tr.trace_logp[(size_t)c * tr.trace_iters + row] = tr.trace_logp[(size_t)src * tr.trace_iters + (row - 1)];
    if (tr.decisions) tr.decisions[(size_t)c * tr.trace_iters + row] = src != c ? DREAMZS_DECISION_SWAPPED : 0u;
    if (src != c) {
      last_like[c] = c == a ? ws[6] : ws[4];
      last_prior[c] = c == a ? ws[7] : ws[5];
    }
  }
}

}  // namespace dreamzs

using namespace dreamzs;

extern "C" int dreamzs_pt_swap(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr, int64_t iter,
                               const double *temperature, double *swap_ws, void *stream) {
  if (!cfg || !st || !tr || cfg->abi_version != DREAMZS_ABI_VERSION || !temperature || !swap_ws || iter < 0) return DREAMZS_E_BADARG;
  if (cfg->nchains_local != cfg->nchains_global || cfg->chain_begin != 0) return DREAMZS_E_UNSUPPORTED;   // the pair may be any two chains
  if (cfg->nchains_global < 2 || cfg->ld < cfg->ndim || (cfg->ld & 3)) return DREAMZS_E_BADARG;
  if (!st->X || !st->last_like || !st->last_prior || !tr->trace || !tr->trace_logp) return DREAMZS_E_BADARG;
  if (tr->trace_offset < 1 || tr->trace_offset >= tr->trace_iters) return DREAMZS_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  pt_decide_kernel<<<1, 32, 0, s>>>(*cfg, st->last_like, st->last_prior, temperature, iter, swap_ws);
  pt_apply_kernel<<<(cfg->nchains_local + 31) / 32, 256, 0, s>>>(*cfg, st->X, st->last_like, st->last_prior, *tr, swap_ws);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}
