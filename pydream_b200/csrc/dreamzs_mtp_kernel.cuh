// Point-parallel multi-try kernels (MT-DREAM(ZS), pydream/Dream.py:270-323, 839-917), sm_100a.
//
// The generic kernel (dreamzs_step_kernel.cuh) gives a chain ONE lane-group and walks the 2k-1 points of a multi-try
// iteration one after the other: at the C3 shape (d = 10, k = 5, 4096 chains) that is 4 lanes per chain, 3.5 warps per SM
// and nine dependent rounds of draws + archive gathers + log-density per iteration.  Here a chain owns k lane-groups of G
// lanes, one per point of a batch (a lane holds R chunks of 4 dimensions), and a warp holds as many chains as fit:
// d = 10, k = 5 -> G = 2, R = 2: ten lanes per chain, three chains per warp.  The k proposals are assembled, bounded and
// evaluated side by side, then -- around the selected one -- the k-1 reference points: two rounds instead of 2k-1.  Every
// variate is a function of (seed, chain, iteration, stream, call number, block), so the points need no ordering among
// themselves except for the data-dependent rand() calls of the boundary redraws, whose call numbers are an exclusive
// prefix over the points (bounds_pp below).  All synchronisation is per chain (MtLay::cmask): the chains of a warp
// diverge freely (snooker / DE iterations, regenerated batches).
//
// Draws, arithmetic and decisions are those of dreamzs_step_kernel<G, R, true> / the oracle (the device functions are
// shared); the chain state is replicated in every lane-group.
#pragma once
#include "dreamzs_step_kernel.cuh"

namespace dreamzs {

// where a lane sits: its chain's lanes (mask, first lane), its lane within the chain, its point
struct MtLay { unsigned cmask; int cbase, cl, pt; };

// Boundary handling of one batch, points side by side (pydream/Dream.py:734-791).  The sequential form (apply_bounds)
// makes one rand() call per point and per non-empty pass (lower set, then upper set) in point order: point p's calls
// start after those of the points before it.
template <int G, int R>
__device__ __forceinline__ void bounds_redraw(const Ctx<G, R> &c, const Stream &s, uint32_t call, unsigned m, double (&p)[R][4]) {
  const StepParams &P = c.P;
  int before_round = 0;  // out-of-bounds dims in earlier rounds (all lanes): ranks run in dimension order = chunk order
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int mine = __popc((m >> (4 * r)) & 15u);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
      const int t = __shfl_up_sync(c.gmask, incl, o, G);
      if (c.g >= o) incl += t;
    }
    int rank = before_round + incl - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((m >> (4 * r + j)) & 1u) {
        const int i = c.dim0(r) + j;
        const uint4 w = s.block(call, ST_RAND, (uint32_t)(rank >> 2));
        const uint32_t ww = (rank & 3) == 0 ? w.x : (rank & 3) == 1 ? w.y : (rank & 3) == 2 ? w.z : w.w;
        const double mn = P.st.mins[i], mx = P.st.maxs[i];
        p[r][j] = mn + u32_of(ww) * (mx - mn);
        ++rank;
      }
    before_round += gsum_int<G>(mine, c.gmask);
  }
}
template <int G, int R>
__device__ __forceinline__ void bounds_pp(const Ctx<G, R> &c, const MtLay &L, Stream &s, bool act, int n, double (&p)[R][4]) {
  const StepParams &P = c.P;
  unsigned lo = 0, hi = 0;
  int tl = 0, th = 0;
  if (act) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = c.dim0(r) + j;
        if (i < c.d) {
          const double mn = P.st.mins[i], mx = P.st.maxs[i];
          double v = p[r][j];
          if (v < mn) v = 2 * mn - v;
          else if (v > mx) v = 2 * mx - v;
          p[r][j] = v;
          if (v < mn) lo |= 1u << (4 * r + j);
          if (v > mx) hi |= 1u << (4 * r + j);
        }
      }
    tl = gsum_int<G>(__popc(lo), c.gmask);
    th = gsum_int<G>(__popc(hi), c.gmask);
  }
  const int cnt = (tl > 0) + (th > 0);
  int base = 0, total = 0;
  for (int q = 0; q < n; ++q) {
    const int cq = __shfl_sync(L.cmask, cnt, L.cbase + q * G);
    if (q < L.pt) base += cq;
    total += cq;
  }
  if (total == 0) return;
  if (act) {
    if (tl > 0) bounds_redraw<G, R>(c, s, s.n_rand + (uint32_t)base, lo, p);
    if (th > 0) bounds_redraw<G, R>(c, s, s.n_rand + (uint32_t)base + (tl > 0 ? 1u : 0u), hi, p);
  }
  s.n_rand += (uint32_t)total;
}

// One batch of n points around `ctr` (generate_proposal_points, Dream.py:628-732): lane-group pt makes point pt, leaves it
// in `out` / its slot and its scalars in pri / lik / snk [pt].  Returns np.any(gamma == 1.0) of the batch.
template <int G, int R>
__device__ __noinline__ bool mtp_batch(const Ctx<G, R> &c, const MtLay L, Stream &s, const Decisions &dc, int n, int64_t M,
                                       const double (&ctr)[R][4], double *slot, double (&out)[R][4], double *pri,
                                       double *lik, double *snk) {
  const StepParams &P = c.P;
  const Bases b = {s.n_sample, s.n_normal, s.n_uvec};
  bool gamma_one = false;
  double gamma = 0.0;
  const uint32_t m_base = s.n_multinomial;
  if (dc.run_snooker) {
    (void)multinomial2(s, P.cfg.p_gamma_unity);              // drawn and discarded (Dream.py:615-618)
    gamma = 1.2 + (2.2 - 1.2) * uniform_scalar(s);
    if (gamma == 1.0) gamma_one = true;
  }
  const int pt = L.pt;
  const bool act = pt < n;
  double sl = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) out[r][0] = out[r][1] = out[r][2] = out[r][3] = 0.0;
  if (act) {
    if (dc.run_snooker) {
      double D;
      snooker_point<G, R>(c, s, b, n, pt, M, gamma, ctr, out, sl, D);
    } else {
      s.n_multinomial = m_base + pt;                         // gamma-unity draw of point pt
      de_point<G, R>(c, s, dc, b, n, pt, M, ctr, out, gamma_one);
    }
  }
  if (P.cfg.hardboundaries && !P.all_flat) bounds_pp<G, R>(c, L, s, act, n, out);
  if (act) {
    store_slot<G, R>(c, slot, out);
    __syncwarp(c.gmask);
    double pr, lk;
    eval_logp<G, R>(c, out, slot, pr, lk);
    if (c.g == 0) { pri[pt] = pr; lik[pt] = lk; snk[pt] = sl; }
  }
  gamma_one = __any_sync(L.cmask, gamma_one);
  if (dc.run_snooker) s.n_sample = b.s + 3 * n;
  else { s.n_sample = b.s + n; s.n_normal = b.n + n; s.n_uvec = b.u + 2 * n; s.n_multinomial = m_base + n; }
  __syncwarp(L.cmask);
  return gamma_one;
}

// What one multi-try iteration decides (the state update itself is mt_commit)
struct MtOutcome { int sel; bool gamma_one, accepted; double new_prior, new_like; };

// The multi-try selection and acceptance arithmetic is a handful of exp / log / divide on k (or 2k) numbers.  Every lane
// of the chain would compute all of them -- and on B200 a warp-wide fp64 instruction costs the same two issue cycles of
// the one fp64 pipe whether one lane needs it or 32 -- so lane l of the chain computes the l-th exponential only and the
// values are exchanged by shuffles; the sums run in the reference's order on every lane (same bits everywhere).
// (A chain has k G >= 2k lanes for G >= 2; with G = 1 the lanes past the chain's compute nothing and the second half of
// the exponentials of the ratio takes a second pass.)
__device__ __forceinline__ double shfl_d(unsigned mask, double v, int src) { return __shfl_sync(mask, v, src); }

// mt_choose_proposal_pt (Dream.py:883-917) given the uniform of the multinomial draw: inverse CDF on a running sum
__device__ __forceinline__ int mt_choose(const MtLay &L, const double *pri, const double *lik, double Tc, int k, double u) {
  double mx = Tc * lik[0] + pri[0];
  for (int p = 1; p < k; ++p) { const double v = Tc * lik[p] + pri[p]; if (v > mx) mx = v; }
  const int lp = L.cl < k ? L.cl : 0;
  const double e = exp((Tc * lik[lp] + pri[lp]) - mx);        // lane p of the chain: exp(log_ps[p] - max)
  double sum = 0.0;
  for (int p = 0; p < k; ++p) { const double ep = shfl_d(L.cmask, e, L.cbase + p); sum = (p == 0) ? ep : sum + ep; }
  const double pr = e / sum;                                   // lane p: the probability of point p
  double acc = 0.0;
  int idx = k - 1;
  bool found = false;
  for (int p = 0; p < k; ++p) {
    acc = acc + shfl_d(L.cmask, pr, L.cbase + p);
    if (!found && u < acc) { idx = p; found = true; }
  }
  return idx;
}

// log of the multi-try acceptance ratio, Dream.py:304-323 (reference point k-1 is the current state); the chain has at
// least 2k lanes (G >= 2)
__device__ __forceinline__ double mt_ratio(const MtLay &L, const double *pri, const double *lik, const double *snk, const double *rpri,
                                           const double *rlik, const double *rsnk, double last_prior, double last_like,
                                           double Tc, int k, bool run_snooker) {
  // lane p < k of the chain holds the proposal term of point p, lane k + p its reference term
  const int p = L.cl < k ? L.cl : (L.cl < 2 * k ? L.cl - k : 0);
  const double lps = Tc * lik[p] + pri[p];
  const double rl = (p == k - 1) ? last_like : rlik[p], rp = (p == k - 1) ? last_prior : rpri[p];
  const double rlps = Tc * rl + rp;
  double tpv, trv;
  if (run_snooker) {                                                                 // Dream.py:306-313
    const double rs = (p == k - 1) ? 0.0 : rsnk[p];
    tpv = lps + snk[p]; trv = rlps + rs + snk[p];
  } else { tpv = lps; trv = rlps; }
  double m2 = shfl_d(L.cmask, tpv, L.cbase);
  for (int q = 0; q < k; ++q) {
    const double a = shfl_d(L.cmask, tpv, L.cbase + q), b = shfl_d(L.cmask, trv, L.cbase + q);
    if (a > m2) m2 = a;
    if (b > m2) m2 = b;
  }
  const double ex = exp((L.cl < k ? tpv : trv) - m2);
  double swp = 0.0, swr = 0.0;
  for (int q = 0; q < k; ++q) {
    const double a = shfl_d(L.cmask, ex, L.cbase + q), b = shfl_d(L.cmask, ex, L.cbase + k + q);
    swp = q == 0 ? a : swp + a; swr = q == 0 ? b : swr + b;
  }
  return nan_to_num(log(swp / swr));                                                 // Dream.py:320-323
}

// One whole multi-try iteration from the random stream (Dream.py:193-362 with multitry > 1): decisions, the k proposals
// (regenerated while none has a finite log-posterior), the choice, the reference set, the acceptance test.  q = the
// selected proposal.  Out of line: the fused kernel's loop body, and the two-stage kernel's way out when a batch has to be
// regenerated (the draws made ahead by the draw kernel no longer apply then).
template <int G, int R>
__device__ __noinline__ void mtp_iteration(const Ctx<G, R> &c, const MtLay L, int64_t iter, uint32_t c_global, double *slot,
                                           const double (&x0)[R][4], double last_prior, double last_like, double Tc,
                                           int64_t M, Decisions &dc, double (&q)[R][4], MtOutcome &o) {
  const StepParams &P = c.P;
  const int k = P.cfg.multitry, ld = c.ld;
  double *pri = c.scal, *lik = c.scal + DREAMZS_MAX_MULTITRY, *snk = c.scal + 2 * DREAMZS_MAX_MULTITRY;
  double *rpri = pri + k, *rlik = lik + k, *rsnk = snk + k;
  Stream s; s.init(P.cfg.seed, c_global, (uint32_t)iter);
  dc.run_snooker = 0;
  if (P.cfg.snooker != 0) dc.run_snooker = multinomial2(s, P.cfg.snooker) == 0;      // set_snooker, Dream.py:542-554
  dc.cr_idx = multinomial_index(s, P.st.cr_probs, P.cfg.nCR);                        // set_CR, :556-569
  dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
  dc.delta = 1;
  if (P.cfg.nDEpairs > 1) {                                                          // set_DEpair, :571-583
    const uint4 w = s.block(s.n_randint++, ST_RANDINT, 0);
    dc.delta = 1 + (int)(((uint64_t)w.x * (uint64_t)P.cfg.nDEpairs) >> 32);
  }
  dc.lvl_idx = multinomial_index(s, P.st.gamma_probs, P.cfg.ngamma);                 // set_gamma_level, :585-599
  double pp[R][4];
  for (int guard = 0;; ++guard) {                                                    // Dream.py:278-289
    o.gamma_one = mtp_batch<G, R>(c, L, s, dc, k, M, x0, slot, pp, pri, lik, snk);
    bool anyfinite = false;
    for (int p = 0; p < k; ++p) anyfinite |= isfinite(Tc * lik[p] + pri[p]);
    if (anyfinite || guard >= 1000) break;
  }
  {
    const uint4 w = s.block(s.n_multinomial++, ST_MULTINOMIAL, 0);
    o.sel = mt_choose(L, pri, lik, Tc, k, u53_of(w.x, w.y));
  }
  load_slot<G, R>(c, c.slots + (size_t)o.sel * ld, q);
  o.new_prior = pri[o.sel]; o.new_like = lik[o.sel];
  __syncwarp(L.cmask);   // every lane-group has read the selected point before the reference set overwrites the slots
  o.gamma_one = mtp_batch<G, R>(c, L, s, dc, k - 1, M, q, slot, pp, rpri, rlik, rsnk);   // reference set, Dream.py:295-303
  const double mr = mt_ratio(L, pri, lik, snk, rpri, rlik, rsnk, last_prior, last_like, Tc, k, dc.run_snooker != 0);
  o.accepted = false;
  if (isfinite(mr)) o.accepted = log(uniform_scalar(s)) < mr;                        // metrop_select, :980-998
}

// State update (Dream.py:336-347: "accepted" is inferred from the state having changed), trace row (core.py:114-115)
// and archive append (record_history, Dream.py:360-362, 919-938) of one iteration; lane-group 0 of the chain writes.
template <int G, int R>
__device__ __forceinline__ void mt_commit(const Ctx<G, R> &c, const MtLay &L, const StepParams &P, int it, int64_t iter, int c_local,
                                          uint32_t c_global, int64_t M, const Decisions &dc, const MtOutcome &o,
                                          const double (&q)[R][4], double (&x0)[R][4], double &last_prior, double &last_like,
                                          double Tc) {
  int changed = 0;
  if (o.accepted) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) { changed |= (q[r][j] != x0[r][j]); x0[r][j] = q[r][j]; }
  }
  changed = gsum_int<G>(changed, c.gmask) != 0;
  if (changed) { last_prior = o.new_prior; last_like = o.new_like; }
  const int64_t trow = P.tr.trace_offset + it;
  if (L.pt == 0) {
    store_row<G, R>(c, P.tr.trace + ((size_t)c_local * P.tr.trace_iters + trow) * c.ld, x0);
    if (c.g == 0) {
      P.tr.trace_logp[(size_t)c_local * P.tr.trace_iters + trow] = Tc * last_like + last_prior;   // core.py:115 (T = 1), :176
      if (P.tr.decisions)
        P.tr.decisions[(size_t)c_local * P.tr.trace_iters + trow] =
            pack_decision(changed, dc.run_snooker, dc.cr_idx, dc.lvl_idx, dc.delta, o.sel, o.gamma_one, o.accepted);
    }
    if (iter % P.cfg.history_thin == 0) {   // only the last iteration of a launch may append
      store_row<G, R>(c, P.st.Z + (size_t)(M + c_global) * c.ld, x0);
      for (int pz = 0; pz < P.npeers; ++pz) store_row<G, R>(c, P.peer_Z[pz] + (size_t)(M + c_global) * c.ld, x0);   // replicas over NVLink
      if (P.publish_k) {
        __threadfence_system();
        __syncwarp(c.gmask);
        if (c.g == 0) peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
      }
    }
  }
  __syncwarp(L.cmask);
}

#ifndef DZ_MTP_MINBLOCKS
#define DZ_MTP_MINBLOCKS 3
#endif

// what the multi-try kernels share at entry: the target table staged in shared memory, the lane's place, the chain's
// context and state
#define DZ_MT_PROLOGUE()                                                                                              \
  extern __shared__ __align__(16) double smem[];                                                                     \
  const int d = P.cfg.ndim, ld = P.cfg.ld, k = P.cfg.multitry;                                                        \
  const double *table = P.st.target_table;                                                                           \
  double *sm_chain = smem;                                                                                           \
  if (P.table_in_smem) {                                                                                             \
    for (int i = threadIdx.x; i < P.table_doubles; i += blockDim.x) smem[i] = table[i];                              \
    table = smem;                                                                                                    \
    sm_chain = smem + ((P.table_doubles + 1) & ~1);                                                                  \
    __syncthreads();                                                                                                 \
  }                                                                                                                  \
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;                                                        \
  const int LC = k * G, CPW = 32 / LC;                 /* lanes per chain, chains per warp */                        \
  const int cw = lane / LC;                                                                                          \
  const int chain_in_cta = warp * CPW + cw;                                                                          \
  const int c_local = (blockIdx.x * (blockDim.x >> 5) + warp) * CPW + cw;                                            \
  if (cw >= CPW || c_local >= P.cfg.nchains_local) return;      /* whole chains leave together */                    \
  MtLay L;                                                                                                           \
  L.cbase = cw * LC; L.cl = lane - L.cbase; L.pt = L.cl / G;                                                          \
  L.cmask = (LC == 32 ? 0xffffffffu : ((1u << LC) - 1u)) << L.cbase;                                                  \
  const int per_chain = k * ld + 3 * DREAMZS_MAX_MULTITRY;                                                           \
  Ctx<G, R> c{P, table, sm_chain + (size_t)chain_in_cta * per_chain, nullptr,                                        \
              ((1u << G) - 1u) << (L.cbase + L.pt * G), L.cl - L.pt * G, d, ld};                                      \
  c.scal = c.slots + (size_t)k * ld;                                                                                 \
  double *slot = c.slots + (size_t)L.pt * ld;                                                                        \
  const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);                                                 \
  double x0[R][4];                                                                                                   \
  load_slot<G, R>(c, P.st.X + (size_t)c_local * ld, x0);                                                             \
  double last_prior = P.st.last_prior[c_local], last_like = P.st.last_like[c_local];                                 \
  const double Tc = P.temperature ? P.temperature[c_local] : 1.0;                                                    \
  const int64_t M = P.archive_rows

// Fused form: draws and evaluation in one kernel (used when no draw scratch is given).
// (__grid_constant__: the out-of-line functions take the parameter block by reference; without it the kernel would copy
// all of it to local memory first)
template <int G, int R>
__global__ void __launch_bounds__(128, DZ_MTP_MINBLOCKS) dreamzs_mtp_kernel(const __grid_constant__ StepParams P) {
  DZ_MT_PROLOGUE();
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // an earlier wait timed out: the host raises
  if (P.wait_k) {   // sharded archive: the peers' rows of the previous append must have landed in this replica
    int ok = 1;
    if (L.cl == 0) ok = peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error) ? 1 : 0;
    if (!__shfl_sync(L.cmask, ok, L.cbase)) return;   // timed out: states stay as they are
  }
#pragma unroll 1
  for (int it = 0; it < P.niter; ++it) {
    const int64_t iter = P.iter_begin + it;
    Decisions dc;
    MtOutcome o;
    double q[R][4];
    mtp_iteration<G, R>(c, L, iter, c_global, slot, x0, last_prior, last_like, Tc, M, dc, q, o);
    mt_commit<G, R>(c, L, P, it, iter, c_local, c_global, M, dc, o, q, x0, last_prior, last_like, Tc);
  }
  if (L.pt == 0) {
    store_row<G, R>(c, P.st.X + (size_t)c_local * ld, x0);
    if (c.g == 0) { P.st.last_prior[c_local] = last_prior; P.st.last_like[c_local] = last_like; }
  }
}

// ================================================================ two-stage form: draw kernel + chain kernel
// Everything a multi-try iteration draws is a function of the Philox counters and of the archive, not of the chain
// state (Dream.py:628-732: decisions, archive rows, e, zeta, the crossover mask and hence d' and gamma), as long as no
// batch is regenerated.  dreamzs_mtdraw_kernel makes all of it for every (chain, iteration, point) of the window in
// parallel -- dense lanes, thousands of independent Philox streams, the gathers of a whole window in flight -- and
// leaves per (chain, iteration) a record in a scratch buffer (L2-resident at the benchmark shapes):
//   [0] uniform of the selection draw   [1] log of the Metropolis uniform   [2], [3] snooker gamma of the two batches
//   [4] two uint32: decisions (bit 0 snooker, 1-4 CR index, 5-8 gamma level, 9-12 DE pairs), gamma == 1 bits per point
//   [8 + 2 p ld ...) point p (proposals 0..k-1, reference points k..2k-2): DE: J = (e*gamma)*diff and zeta, both 0 where
//                    the crossover keeps the centre;  snooker: z and z1 - z2
// dreamzs_mtchain_kernel then walks the Markov chains: proposal = (centre + J) + zeta, bounds, log-density, choice,
// reference set, acceptance -- no random numbers except the boundary redraws.  A batch without a finite log-posterior
// (Dream.py:282-289 regenerates it, which shifts every later call number) sends that iteration through mtp_iteration.
__host__ __device__ inline int mt2_record_doubles(int k, int ld) { return 8 + (2 * k - 1) * 2 * ld; }

// scalar draws of every (chain, iteration) pair of the window + the assembled decisions
static __global__ void __launch_bounds__(256) dreamzs_mtdraw_scalars_kernel(const __grid_constant__ StepParams P) {
  constexpr int CT = 64;                           // pairs of a CTA
  __shared__ uint2 scr[9 * CT];
  const int k = P.cfg.multitry, wn = P.niter;
  const int ncts = P.cfg.nchains_local * wn;
  const int ct0 = blockIdx.x * CT;
  const int nct = min(CT, ncts - ct0);
  const int S = mt2_record_doubles(k, P.cfg.ld);
  const uint32_t k0 = (uint32_t)P.cfg.seed, k1 = (uint32_t)(P.cfg.seed >> 32);
  const uint32_t s0 = P.cfg.snooker != 0 ? 1u : 0u, m0 = s0 + 2u;
  const int tid = threadIdx.x;
  // one Philox block per (kind, pair):
  //   0 snooker, 1 CR, 2 gamma level (multinomial calls 0, s0, s0+1), 3 DE pairs (randint), 4-6 np.random.uniform()
  //   calls 0-2, 7 / 8 the selection multinomial after DE / snooker batches (calls m0 + k / m0 + 1)
  for (int task = tid; task < 9 * nct; task += 256) {
    const int kind = task / nct, cti = task - kind * nct;
    const int ct = ct0 + cti, c_local = ct / wn, itb = ct - c_local * wn;
    uint32_t call = 0, st = ST_MULTINOMIAL;
    if (kind == 1) call = s0;
    else if (kind == 2) call = s0 + 1;
    else if (kind == 3) st = ST_RANDINT;
    else if (kind >= 4 && kind <= 6) { st = ST_UNIFORM_SCAL; call = (uint32_t)(kind - 4); }
    else if (kind == 7) call = m0 + (uint32_t)k;
    else if (kind == 8) call = m0 + 1;
    const uint4 w = philox4x32(0u, (call << 3) | st, (uint32_t)(P.iter_begin + itb), (uint32_t)(P.cfg.chain_begin + c_local), k0, k1);
    scr[kind * CT + cti] = make_uint2(w.x, w.y);
  }
  __syncthreads();
  if (tid < nct) {
    const int cti = tid, ct = ct0 + cti;
    auto u53 = [&](int kind) { const uint2 w = scr[kind * CT + cti]; return u53_of(w.x, w.y); };
    auto invcdf = [&](double u, const double *p, int n) {
      double acc = 0.0;
      int idx = n - 1;
      bool found = false;
      for (int j = 0; j < n; ++j) {
        acc = acc + p[j];
        if (!found && u < acc) { idx = j; found = true; }
      }
      return idx;
    };
    const bool snk = s0 != 0u && u53(0) < 0.0 + P.cfg.snooker;
    const int cr = invcdf(u53(1), P.st.cr_probs, P.cfg.nCR), lvl = invcdf(u53(2), P.st.gamma_probs, P.cfg.ngamma);
    int delta = 1;
    if (P.cfg.nDEpairs > 1) delta = 1 + (int)(((uint64_t)scr[3 * CT + cti].x * (uint64_t)P.cfg.nDEpairs) >> 32);
    const uint32_t dec = (snk ? 1u : 0u) | ((uint32_t)cr << 1) | ((uint32_t)lvl << 5) | ((uint32_t)delta << 9);
    double *rec = P.st.draw_ws + (size_t)ct * S;
    rec[0] = u53(snk ? 8 : 7);
    rec[1] = log(u53(snk ? 6 : 4));                      // the Metropolis uniform is np.random.uniform() call 2 / 0
    rec[2] = 1.2 + (2.2 - 1.2) * u53(4);                 // snooker gamma, Dream.py:618
    rec[3] = 1.2 + (2.2 - 1.2) * u53(5);
    *reinterpret_cast<uint2 *>(rec + 4) = make_uint2(dec, 0u);   // the gamma == 1 bits are set by the points' kernel
  }
}

// the points: one lane-group per (pair, point), flat over the window
template <int G>
__global__ void __launch_bounds__(256, 4) dreamzs_mtdraw_kernel(const __grid_constant__ StepParams P) {
  const int d = P.cfg.ndim, ld = P.cfg.ld, k = P.cfg.multitry, npts = 2 * k - 1, wn = P.niter;
  const int S = mt2_record_doubles(k, ld);
  const uint32_t s0 = P.cfg.snooker != 0 ? 1u : 0u, m0 = s0 + 2u;
  const int tid = threadIdx.x;
  if (P.wait_k) {   // sharded archive: the peers' rows of the previous append must have landed in this replica
    if (tid == 0) peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error);
    __syncthreads();
  }
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // timed out (now or earlier): the chain kernel leaves too
  const int lane = tid & 31, g = lane & (G - 1);
  Ctx<G, 1> c{P, nullptr, nullptr, nullptr, G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1))), g, d, ld};
  const int i0 = 4 * g;
  const int64_t unit = (int64_t)blockIdx.x * (256 / G) + tid / G;
  if (unit >= (int64_t)P.cfg.nchains_local * wn * npts) return;     // (whole lane-groups leave together)
  const int ct = (int)(unit / npts), p = (int)(unit - (int64_t)ct * npts);
  const int c_local = ct / wn, itb = ct - c_local * wn;
  double *rec = P.st.draw_ws + (size_t)ct * S;
  const uint32_t dec = *reinterpret_cast<const uint32_t *>(rec + 4);
  Decisions dc;
  dc.run_snooker = (int)(dec & 1u); dc.cr_idx = (int)((dec >> 1) & 15u); dc.lvl_idx = (int)((dec >> 5) & 15u);
  dc.delta = (int)((dec >> 9) & 15u);
  dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
  Stream s; s.init(P.cfg.seed, (uint32_t)(P.cfg.chain_begin + c_local), (uint32_t)(P.iter_begin + itb));
  const bool second = p >= k;
  const int n = second ? k - 1 : k, pp = second ? p - k : p;
  double A[1][4], B[1][4];
  if (dc.run_snooker) {
    const Bases b = {second ? 3u * (uint32_t)k : 0u, 0u, 0u};
    snooker_rows<G, 1>(c, s, b, n, pp, P.archive_rows, A, B);
  } else {
    // call numbers of the batch: samples / normals from b.s = b.n, uniforms from b.u, the gamma-unity multinomial
    const Bases b = {second ? (uint32_t)k : 0u, second ? (uint32_t)k : 0u, second ? 2u * (uint32_t)k : 0u};
    s.n_multinomial = (second ? m0 + (uint32_t)k + 1u : m0) + (uint32_t)pp;
    unsigned reset;
    bool gone = false;
    de_draw<G, 1>(c, s, dc, b, n, pp, P.archive_rows, A, B, reset, gone);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (((reset >> j) & 1u) || i0 + j >= d) { A[0][j] = 0.0; B[0][j] = 0.0; }
    if (gone && g == 0) atomicOr(reinterpret_cast<unsigned int *>(rec + 4) + 1, 1u << p);
  }
  if (i0 < ld) {
    double *rp = rec + 8 + (size_t)p * 2 * ld + i0;
    *reinterpret_cast<double2 *>(rp) = make_double2(A[0][0], A[0][1]);
    *reinterpret_cast<double2 *>(rp + 2) = make_double2(A[0][2], A[0][3]);
    *reinterpret_cast<double2 *>(rp + ld) = make_double2(B[0][0], B[0][1]);
    *reinterpret_cast<double2 *>(rp + ld + 2) = make_double2(B[0][2], B[0][3]);
  }
}

// one batch of the chain kernel: lane-group pt assembles point pt from the record, bounds, log-density
template <int G, int R>
__device__ __forceinline__ void mt2_batch(const Ctx<G, R> &c, const MtLay &L, Stream &s, bool run_snooker, double gamma, int n,
                                          const double (&ctr)[R][4], const double (&A)[R][4], double (&B)[R][4], double *slot,
                                          double *pri, double *lik, double *snk) {
  const StepParams &P = c.P;
  const bool act = L.pt < n;
  double out[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) out[r][0] = out[r][1] = out[r][2] = out[r][3] = 0.0;
  double sl = 0.0;
  if (act) {
    if (run_snooker) {
      double D;
      snooker_compute<G, R>(c, n, gamma, ctr, A, B, out, sl, D);
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[r][j] = ctr[r][j] + A[r][j] + B[r][j];   // Dream.py:717 (centre kept where J = zeta = 0)
    }
  }
  if (P.cfg.hardboundaries && !P.all_flat) bounds_pp<G, R>(c, L, s, act, n, out);
  if (act) {
    store_slot<G, R>(c, slot, out);
    __syncwarp(c.gmask);
    double pr, lk;
    eval_logp<G, R>(c, out, slot, pr, lk);
    if (c.g == 0) { pri[L.pt] = pr; lik[L.pt] = lk; snk[L.pt] = sl; }
  }
  __syncwarp(L.cmask);
}

template <int G, int R>
__global__ void __launch_bounds__(128, DZ_MTP_MINBLOCKS) dreamzs_mtchain_kernel(const __grid_constant__ StepParams P) {
  DZ_MT_PROLOGUE();
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // the draw kernel's wait for the peers timed out
  double *pri = c.scal, *lik = c.scal + DREAMZS_MAX_MULTITRY, *snk = c.scal + 2 * DREAMZS_MAX_MULTITRY;
  double *rpri = pri + k, *rlik = lik + k, *rsnk = snk + k;
  const int S = mt2_record_doubles(k, ld);
  const int pt = L.pt;
  const double *rec = P.st.draw_ws + (size_t)c_local * P.niter * S;
#pragma unroll 1
  for (int it = 0; it < P.niter; ++it, rec += S) {
    const int64_t iter = P.iter_begin + it;
    // the record: scalars, this lane-group's proposal and reference point (neither depends on the chain state)
    if (it + 1 < P.niter)   // the next record: on its way to L1 while this iteration runs
      for (int o = L.cl * 16; o < S; o += LC * 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + S + o));
    const double u_sel = rec[0], logu = rec[1], g1 = rec[2], g2 = rec[3];
    const uint2 wd = *reinterpret_cast<const uint2 *>(rec + 4);
    double A1[R][4], B1[R][4], A2[R][4], B2[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = c.dim0(r);
#pragma unroll
      for (int j = 0; j < 4; ++j) { A1[r][j] = 0.0; B1[r][j] = 0.0; A2[r][j] = 0.0; B2[r][j] = 0.0; }
      if (i0 < ld) {
        if (pt < k) {
          const double *r1 = rec + 8 + (size_t)pt * 2 * ld + i0;
          const double2 a = *reinterpret_cast<const double2 *>(r1), b = *reinterpret_cast<const double2 *>(r1 + 2);
          const double2 e = *reinterpret_cast<const double2 *>(r1 + ld), f = *reinterpret_cast<const double2 *>(r1 + ld + 2);
          A1[r][0] = a.x; A1[r][1] = a.y; A1[r][2] = b.x; A1[r][3] = b.y; B1[r][0] = e.x; B1[r][1] = e.y; B1[r][2] = f.x; B1[r][3] = f.y;
        }
        if (pt < k - 1) {
          const double *r2 = rec + 8 + (size_t)(k + pt) * 2 * ld + i0;
          const double2 a = *reinterpret_cast<const double2 *>(r2), b = *reinterpret_cast<const double2 *>(r2 + 2);
          const double2 e = *reinterpret_cast<const double2 *>(r2 + ld), f = *reinterpret_cast<const double2 *>(r2 + ld + 2);
          A2[r][0] = a.x; A2[r][1] = a.y; A2[r][2] = b.x; A2[r][3] = b.y; B2[r][0] = e.x; B2[r][1] = e.y; B2[r][2] = f.x; B2[r][3] = f.y;
        }
      }
    }
    Decisions dc;
    dc.run_snooker = (int)(wd.x & 1u); dc.cr_idx = (int)((wd.x >> 1) & 15u); dc.lvl_idx = (int)((wd.x >> 5) & 15u);
    dc.delta = (int)((wd.x >> 9) & 15u);
    dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
    Stream s; s.init(P.cfg.seed, c_global, (uint32_t)iter);      // (only the boundary redraws draw here)
    MtOutcome o;
    double q[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) q[r][j] = x0[r][j];
    bool fast = true;
#pragma unroll 1
    for (int bt = 0; bt < 2; ++bt) {   // the proposals around x0, then the reference set around the selected one (one copy of the code)
      if (bt) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j) { A1[r][j] = A2[r][j]; B1[r][j] = B2[r][j]; }
      }
      mt2_batch<G, R>(c, L, s, dc.run_snooker != 0, bt ? g2 : g1, bt ? k - 1 : k, q, A1, B1, slot, bt ? rpri : pri, bt ? rlik : lik,
                      bt ? rsnk : snk);
      if (bt == 0) {
        bool anyfinite = false;
        for (int p = 0; p < k; ++p) anyfinite |= isfinite(Tc * lik[p] + pri[p]);
        if (!anyfinite) { fast = false; break; }
        o.sel = mt_choose(L, pri, lik, Tc, k, u_sel);
        load_slot<G, R>(c, c.slots + (size_t)o.sel * ld, q);
        o.new_prior = pri[o.sel]; o.new_like = lik[o.sel];
        __syncwarp(L.cmask);   // every lane-group has read the selected point before the reference set overwrites the slots
      }
    }
    if (fast) {
      o.gamma_one = dc.run_snooker ? g2 == 1.0 : ((wd.y >> k) & ((1u << (k - 1)) - 1u)) != 0u;
      const double mr = mt_ratio(L, pri, lik, snk, rpri, rlik, rsnk, last_prior, last_like, Tc, k, dc.run_snooker != 0);
      o.accepted = isfinite(mr) && logu < mr;                                        // metrop_select, Dream.py:980-998
    } else {
      __syncwarp(L.cmask);
      mtp_iteration<G, R>(c, L, iter, c_global, slot, x0, last_prior, last_like, Tc, M, dc, q, o);
    }
    mt_commit<G, R>(c, L, P, it, iter, c_local, c_global, M, dc, o, q, x0, last_prior, last_like, Tc);
  }
  if (L.pt == 0) {
    store_row<G, R>(c, P.st.X + (size_t)c_local * ld, x0);
    if (c.g == 0) { P.st.last_prior[c_local] = last_prior; P.st.last_like[c_local] = last_like; }
  }
}

template <int G, int R>
int launch_mtp(const StepParams &P, cudaStream_t stream) {
  const int threads = 128, k = P.cfg.multitry;
  const int chains_per_cta = (threads / 32) * (32 / (k * G));
  const int grid = (P.cfg.nchains_local + chains_per_cta - 1) / chains_per_cta;
  const size_t chain_b = (size_t)chains_per_cta * ((size_t)k * P.cfg.ld + 3 * DREAMZS_MAX_MULTITRY) * sizeof(double);
  const size_t table_b = (size_t)((P.table_doubles + 1) & ~1) * sizeof(double);
  StepParams Q = P;
  Q.table_in_smem = (table_b + chain_b <= 200 * 1024) ? 1 : 0;
  Q.nslots = k;
  const size_t smem = chain_b + (Q.table_in_smem ? table_b : 0);
  const bool two_stage = Q.st.draw_ws != nullptr;
  auto kern = two_stage ? dreamzs_mtchain_kernel<G, R> : dreamzs_mtp_kernel<G, R>;
  if (smem > 48 * 1024) {
    static size_t smem_set[2][64] = {{0}};
    if (ensure_dynamic_smem(kern, smem, smem_set[two_stage ? 1 : 0]) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  }
  if (two_stage) {
    constexpr int GD = G * R;          // the draw kernel's lanes per point: one chunk per lane
    const int64_t ncts = (int64_t)Q.cfg.nchains_local * Q.niter, units = ncts * (2 * k - 1);
    dreamzs_mtdraw_scalars_kernel<<<(unsigned)((ncts + 63) / 64), 256, 0, stream>>>(Q);
    dreamzs_mtdraw_kernel<GD><<<(unsigned)((units + 256 / GD - 1) / (256 / GD)), 256, 0, stream>>>(Q);
    if (cudaGetLastError() != cudaSuccess) return DREAMZS_E_LAUNCH;
  }
  kern<<<grid, threads, smem, stream>>>(Q);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
