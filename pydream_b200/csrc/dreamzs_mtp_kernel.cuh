// Point-parallel multi-try kernel (MT-DREAM(ZS), pydream/Dream.py:270-323, 839-917), sm_100a.
//
// The generic kernel (dreamzs_step_kernel.cuh) gives a chain ONE lane-group of G lanes and walks the 2k-1 points of a
// multi-try iteration one after the other: at the C3 shape (d = 10, k = 5, 4096 chains) that is 4 lanes per chain, 3.5
// warps per SM and nine dependent rounds of draws + archive gathers + log-density per iteration.  Here a chain owns a
// whole WARP: 32 / G lane-groups, one per point of the batch.  The k proposals are drawn, gathered, bounded and evaluated
// side by side, then -- around the selected one -- the k-1 reference points: two rounds instead of 2k-1, and 32 / G times
// the warps in flight.  Every variate is a function of (seed, chain, iteration, stream, call number, block), so the
// points need no ordering among themselves except for the data-dependent rand() calls of the boundary redraws, whose
// call numbers are an exclusive prefix over the points (bounds_pp below).
//
// Draws, arithmetic and decisions are those of dreamzs_step_kernel<G, 1, true> / the oracle (the device functions are
// shared); the chain state is replicated in every lane-group.
#pragma once
#include "dreamzs_step_kernel.cuh"

namespace dreamzs {

// Boundary handling of one batch, points side by side (pydream/Dream.py:734-791).  The sequential form (apply_bounds)
// makes one rand() call per point and per non-empty pass (lower set, then upper set) in point order: point p's calls
// start after those of the points before it.
template <int G>
__device__ __forceinline__ void bounds_redraw(const Ctx<G, 1> &c, const Stream &s, uint32_t call, unsigned m, double (&p)[1][4]) {
  const StepParams &P = c.P;
  const int mine = __popc(m & 15u);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const int t = __shfl_up_sync(c.gmask, incl, o, G);
    if (c.g >= o) incl += t;
  }
  int rank = incl - mine;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if ((m >> j) & 1u) {
      const int i = c.dim0(0) + j;
      const uint4 w = s.block(call, ST_RAND, (uint32_t)(rank >> 2));
      const uint32_t ww = (rank & 3) == 0 ? w.x : (rank & 3) == 1 ? w.y : (rank & 3) == 2 ? w.z : w.w;
      const double mn = P.st.mins[i], mx = P.st.maxs[i];
      p[0][j] = mn + u32_of(ww) * (mx - mn);
      ++rank;
    }
}
template <int G>
__device__ __forceinline__ void bounds_pp(const Ctx<G, 1> &c, Stream &s, bool act, int n, int pt, double (&p)[1][4]) {
  const StepParams &P = c.P;
  unsigned lo = 0, hi = 0;
  int tl = 0, th = 0;
  if (act) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = c.dim0(0) + j;
      if (i < c.d) {
        const double mn = P.st.mins[i], mx = P.st.maxs[i];
        double v = p[0][j];
        if (v < mn) v = 2 * mn - v;
        else if (v > mx) v = 2 * mx - v;
        p[0][j] = v;
        if (v < mn) lo |= 1u << j;
        if (v > mx) hi |= 1u << j;
      }
    }
    tl = gsum_int<G>(__popc(lo), c.gmask);
    th = gsum_int<G>(__popc(hi), c.gmask);
  }
  const int cnt = (tl > 0) + (th > 0);
  int base = 0, total = 0;
  for (int q = 0; q < n; ++q) {
    const int cq = __shfl_sync(0xffffffffu, cnt, q * G);
    if (q < pt) base += cq;
    total += cq;
  }
  if (total == 0) return;
  if (act) {
    if (tl > 0) bounds_redraw<G>(c, s, s.n_rand + (uint32_t)base, lo, p);
    if (th > 0) bounds_redraw<G>(c, s, s.n_rand + (uint32_t)base + (tl > 0 ? 1u : 0u), hi, p);
  }
  s.n_rand += (uint32_t)total;
}

// One batch of n points around `ctr` (generate_proposal_points, Dream.py:628-732): lane-group pt makes point pt, leaves it
// in `out` / its slot and its scalars in pri / lik / snk [pt].  Returns np.any(gamma == 1.0) of the batch.
template <int G>
__device__ __noinline__ bool mtp_batch(const Ctx<G, 1> &c, Stream &s, const Decisions &dc, int n, int pt, int64_t M,
                                       const double (&ctr)[1][4], double *slot, double (&out)[1][4], double *pri,
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
  const bool act = pt < n;
  double sl = 0.0;
  out[0][0] = out[0][1] = out[0][2] = out[0][3] = 0.0;
  if (act) {
    if (dc.run_snooker) {
      double D;
      snooker_point<G, 1>(c, s, b, n, pt, M, gamma, ctr, out, sl, D);
    } else {
      s.n_multinomial = m_base + pt;                         // gamma-unity draw of point pt
      de_point<G, 1>(c, s, dc, b, n, pt, M, ctr, out, gamma_one);
    }
  }
  if (P.cfg.hardboundaries && !P.all_flat) bounds_pp<G>(c, s, act, n, pt, out);
  if (act) {
    store_slot<G, 1>(c, slot, out);
    __syncwarp(c.gmask);
    double pr, lk;
    eval_logp<G, 1>(c, out, slot, pr, lk);
    if (c.g == 0) { pri[pt] = pr; lik[pt] = lk; snk[pt] = sl; }
  }
  gamma_one = __any_sync(0xffffffffu, gamma_one);
  if (dc.run_snooker) s.n_sample = b.s + 3 * n;
  else { s.n_sample = b.s + n; s.n_normal = b.n + n; s.n_uvec = b.u + 2 * n; s.n_multinomial = m_base + n; }
  __syncwarp();
  return gamma_one;
}

// What one multi-try iteration decides (the state update itself is mt_commit)
struct MtOutcome { int sel; bool gamma_one, accepted; double new_prior, new_like; };

// The multi-try selection and acceptance arithmetic is a handful of exp / log / divide on k (or 2k) numbers.  Every lane
// of the warp would compute all of them -- and on B200 a warp-wide fp64 instruction costs the same two issue cycles of
// the one fp64 pipe whether one lane needs it or 32 -- so lane l computes the l-th exponential only and the values are
// exchanged by shuffles; the sums run in the reference's order on every lane (same bits everywhere).
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// mt_choose_proposal_pt (Dream.py:883-917) given the uniform of the multinomial draw: inverse CDF on a running sum
__device__ __forceinline__ int mt_choose(const double *pri, const double *lik, double Tc, int k, double u, int lane) {
  double mx = Tc * lik[0] + pri[0];
  for (int p = 1; p < k; ++p) { const double v = Tc * lik[p] + pri[p]; if (v > mx) mx = v; }
  const int lp = lane < k ? lane : 0;
  const double e = exp((Tc * lik[lp] + pri[lp]) - mx);        // lane p: exp(log_ps[p] - max)
  double sum = 0.0;
  for (int p = 0; p < k; ++p) { const double ep = shfl_d(e, p); sum = (p == 0) ? ep : sum + ep; }
  const double pr = e / sum;                                   // lane p: the probability of point p
  double acc = 0.0;
  int idx = k - 1;
  bool found = false;
  for (int p = 0; p < k; ++p) {
    acc = acc + shfl_d(pr, p);
    if (!found && u < acc) { idx = p; found = true; }
  }
  return idx;
}

// log of the multi-try acceptance ratio, Dream.py:304-323 (reference point k-1 is the current state); k <= 16 lanes
__device__ __forceinline__ double mt_ratio(const double *pri, const double *lik, const double *snk, const double *rpri,
                                           const double *rlik, const double *rsnk, double last_prior, double last_like,
                                           double Tc, int k, bool run_snooker, int lane) {
  // lane p < k holds the proposal term of point p, lane k + p its reference term
  const int p = lane < k ? lane : (lane < 2 * k ? lane - k : 0);
  const double lps = Tc * lik[p] + pri[p];
  const double rl = (p == k - 1) ? last_like : rlik[p], rp = (p == k - 1) ? last_prior : rpri[p];
  const double rlps = Tc * rl + rp;
  double tpv, trv;
  if (run_snooker) {                                                                 // Dream.py:306-313
    const double rs = (p == k - 1) ? 0.0 : rsnk[p];
    tpv = lps + snk[p]; trv = rlps + rs + snk[p];
  } else { tpv = lps; trv = rlps; }
  double m2 = shfl_d(tpv, 0);
  for (int q = 0; q < k; ++q) {
    const double a = shfl_d(tpv, q), b = shfl_d(trv, q);
    if (a > m2) m2 = a;
    if (b > m2) m2 = b;
  }
  const double ex = exp((lane < k ? tpv : trv) - m2);
  double swp = 0.0, swr = 0.0;
  for (int q = 0; q < k; ++q) {
    const double a = shfl_d(ex, q), b = shfl_d(ex, k + q);
    swp = q == 0 ? a : swp + a; swr = q == 0 ? b : swr + b;
  }
  return nan_to_num(log(swp / swr));                                                 // Dream.py:320-323
}

// One whole multi-try iteration from the random stream (Dream.py:193-362 with multitry > 1): decisions, the k proposals
// (regenerated while none has a finite log-posterior), the choice, the reference set, the acceptance test.  q = the
// selected proposal.  Out of line: the fused kernel's loop body, and the two-stage kernel's way out when a batch has to be
// regenerated (the draws made ahead by the draw kernel no longer apply then).
template <int G>
__device__ __noinline__ void mtp_iteration(const Ctx<G, 1> &c, int64_t iter, uint32_t c_global, int pt, double *slot,
                                           const double (&x0)[1][4], double last_prior, double last_like, double Tc,
                                           int64_t M, Decisions &dc, double (&q)[1][4], MtOutcome &o) {
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
  double pp[1][4];
  for (int guard = 0;; ++guard) {                                                    // Dream.py:278-289
    o.gamma_one = mtp_batch<G>(c, s, dc, k, pt, M, x0, slot, pp, pri, lik, snk);
    bool anyfinite = false;
    for (int p = 0; p < k; ++p) anyfinite |= isfinite(Tc * lik[p] + pri[p]);
    if (anyfinite || guard >= 1000) break;
  }
  {
    const uint4 w = s.block(s.n_multinomial++, ST_MULTINOMIAL, 0);
    o.sel = mt_choose(pri, lik, Tc, k, u53_of(w.x, w.y), (int)(threadIdx.x & 31));
  }
  load_slot<G, 1>(c, c.slots + (size_t)o.sel * ld, q);
  o.new_prior = pri[o.sel]; o.new_like = lik[o.sel];
  __syncwarp();   // every lane-group has read the selected point before the reference set overwrites the slots
  o.gamma_one = mtp_batch<G>(c, s, dc, k - 1, pt, M, q, slot, pp, rpri, rlik, rsnk);   // reference set, Dream.py:295-303
  const double mr = mt_ratio(pri, lik, snk, rpri, rlik, rsnk, last_prior, last_like, Tc, k, dc.run_snooker != 0, (int)(threadIdx.x & 31));
  o.accepted = false;
  if (isfinite(mr)) o.accepted = log(uniform_scalar(s)) < mr;                        // metrop_select, :980-998
}

// State update (Dream.py:336-347: "accepted" is inferred from the state having changed), trace row (core.py:114-115)
// and archive append (record_history, Dream.py:360-362, 919-938) of one iteration; lane-group 0 writes.
template <int G>
__device__ __forceinline__ void mt_commit(const Ctx<G, 1> &c, const StepParams &P, int it, int64_t iter, int c_local,
                                          uint32_t c_global, bool writer, int64_t M, const Decisions &dc, const MtOutcome &o,
                                          const double (&q)[1][4], double (&x0)[1][4], double &last_prior, double &last_like,
                                          double Tc) {
  int changed = 0;
  if (o.accepted) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { changed |= (q[0][j] != x0[0][j]); x0[0][j] = q[0][j]; }
  }
  changed = gsum_int<G>(changed, c.gmask) != 0;
  if (changed) { last_prior = o.new_prior; last_like = o.new_like; }
  const int64_t trow = P.tr.trace_offset + it;
  if (writer) {
    store_row<G, 1>(c, P.tr.trace + ((size_t)c_local * P.tr.trace_iters + trow) * c.ld, x0);
    if (c.g == 0) {
      P.tr.trace_logp[(size_t)c_local * P.tr.trace_iters + trow] = Tc * last_like + last_prior;   // core.py:115 (T = 1), :176
      if (P.tr.decisions)
        P.tr.decisions[(size_t)c_local * P.tr.trace_iters + trow] =
            pack_decision(changed, dc.run_snooker, dc.cr_idx, dc.lvl_idx, dc.delta, o.sel, o.gamma_one, o.accepted);
    }
    if (iter % P.cfg.history_thin == 0) {   // only the last iteration of a launch may append
      store_row<G, 1>(c, P.st.Z + (size_t)(M + c_global) * c.ld, x0);
      for (int pz = 0; pz < P.npeers; ++pz) store_row<G, 1>(c, P.peer_Z[pz] + (size_t)(M + c_global) * c.ld, x0);   // replicas over NVLink
      if (P.publish_k) {
        __threadfence_system();
        __syncwarp(c.gmask);
        if (c.g == 0) peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
      }
    }
  }
  __syncwarp();
}

#ifndef DZ_MTP_MINBLOCKS
#define DZ_MTP_MINBLOCKS 4
#endif

// what the multi-try kernels share at entry: the target table staged in shared memory, the chain's context and state
#define DZ_MT_PROLOGUE()                                                                                              \
  extern __shared__ __align__(16) double smem[];                                                                     \
  constexpr int PP = 32 / G;                      /* points side by side */                                          \
  const int d = P.cfg.ndim, ld = P.cfg.ld;                                                                           \
  const double *table = P.st.target_table;                                                                           \
  double *sm_chain = smem;                                                                                           \
  if (P.table_in_smem) {                                                                                             \
    for (int i = threadIdx.x; i < P.table_doubles; i += blockDim.x) smem[i] = table[i];                              \
    table = smem;                                                                                                    \
    sm_chain = smem + ((P.table_doubles + 1) & ~1);                                                                  \
    __syncthreads();                                                                                                 \
  }                                                                                                                  \
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;                                                        \
  const int c_local = blockIdx.x * (blockDim.x >> 5) + warp;                                                         \
  if (c_local >= P.cfg.nchains_local) return;                                                                        \
  const int pt = lane / G;                                                                                           \
  const int per_chain = PP * ld + 3 * DREAMZS_MAX_MULTITRY;                                                          \
  Ctx<G, 1> c{P, table, sm_chain + (size_t)warp * per_chain, nullptr,                                                \
              G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1))), lane & (G - 1), d, ld};               \
  c.scal = c.slots + (size_t)PP * ld;                                                                                \
  double *slot = c.slots + (size_t)pt * ld;                                                                          \
  const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);                                                 \
  const bool writer = pt == 0;                    /* the lane-group that writes the chain's rows */                  \
  double x0[1][4];                                                                                                   \
  {                                                                                                                  \
    const double *xrow = P.st.X + (size_t)c_local * ld;                                                              \
    const int i0 = c.dim0(0);                                                                                        \
    if (i0 < ld) {                                                                                                   \
      const double2 a = *reinterpret_cast<const double2 *>(xrow + i0), b = *reinterpret_cast<const double2 *>(xrow + i0 + 2); \
      x0[0][0] = a.x; x0[0][1] = a.y; x0[0][2] = b.x; x0[0][3] = b.y;                                                \
    } else x0[0][0] = x0[0][1] = x0[0][2] = x0[0][3] = 0.0;                                                          \
  }                                                                                                                  \
  double last_prior = P.st.last_prior[c_local], last_like = P.st.last_like[c_local];                                 \
  const double Tc = P.temperature ? P.temperature[c_local] : 1.0;                                                    \
  const int64_t M = P.archive_rows

// Fused form: draws and evaluation in one kernel (used when no draw scratch is given).
// (__grid_constant__: the out-of-line functions take the parameter block by reference; without it the kernel would copy
// all of it to local memory first)
template <int G>
__global__ void __launch_bounds__(128, DZ_MTP_MINBLOCKS) dreamzs_mtp_kernel(const __grid_constant__ StepParams P) {
  DZ_MT_PROLOGUE();
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // an earlier wait timed out: the host raises
  if (P.wait_k) {   // sharded archive: the peers' rows of the previous append must have landed in this replica
    int ok = 1;
    if (lane == 0) ok = peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error) ? 1 : 0;
    if (!__shfl_sync(0xffffffffu, ok, 0)) return;   // timed out: states stay as they are
  }
#pragma unroll 1
  for (int it = 0; it < P.niter; ++it) {
    const int64_t iter = P.iter_begin + it;
    Decisions dc;
    MtOutcome o;
    double q[1][4];
    mtp_iteration<G>(c, iter, c_global, pt, slot, x0, last_prior, last_like, Tc, M, dc, q, o);
    mt_commit<G>(c, P, it, iter, c_local, c_global, writer, M, dc, o, q, x0, last_prior, last_like, Tc);
  }
  if (writer) {
    store_row<G, 1>(c, P.st.X + (size_t)c_local * ld, x0);
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
template <int G>
__device__ __forceinline__ void mt2_batch(const Ctx<G, 1> &c, Stream &s, bool run_snooker, double gamma, int n, int pt,
                                          const double (&ctr)[1][4], const double (&A)[1][4], double (&B)[1][4], double *slot,
                                          double *pri, double *lik, double *snk) {
  const StepParams &P = c.P;
  const bool act = pt < n;
  double out[1][4] = {{0.0, 0.0, 0.0, 0.0}};
  double sl = 0.0;
  if (act) {
    if (run_snooker) {
      double D;
      snooker_compute<G, 1>(c, n, gamma, ctr, A, B, out, sl, D);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) out[0][j] = ctr[0][j] + A[0][j] + B[0][j];   // Dream.py:717 (centre kept where J = zeta = 0)
    }
  }
  if (P.cfg.hardboundaries && !P.all_flat) bounds_pp<G>(c, s, act, n, pt, out);
  if (act) {
    store_slot<G, 1>(c, slot, out);
    __syncwarp(c.gmask);
    double pr, lk;
    eval_logp<G, 1>(c, out, slot, pr, lk);
    if (c.g == 0) { pri[pt] = pr; lik[pt] = lk; snk[pt] = sl; }
  }
  __syncwarp();
}

template <int G>
__global__ void __launch_bounds__(128, DZ_MTP_MINBLOCKS) dreamzs_mtchain_kernel(const __grid_constant__ StepParams P) {
  DZ_MT_PROLOGUE();
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // the draw kernel's wait for the peers timed out
  const int k = P.cfg.multitry;
  double *pri = c.scal, *lik = c.scal + DREAMZS_MAX_MULTITRY, *snk = c.scal + 2 * DREAMZS_MAX_MULTITRY;
  double *rpri = pri + k, *rlik = lik + k, *rsnk = snk + k;
  const int S = mt2_record_doubles(k, ld);
  const int i0 = c.dim0(0);
  const double *rec = P.st.draw_ws + (size_t)c_local * P.niter * S;
#pragma unroll 1
  for (int it = 0; it < P.niter; ++it, rec += S) {
    const int64_t iter = P.iter_begin + it;
    // the record: scalars, this lane-group's proposal and reference point (neither depends on the chain state)
    if (it + 1 < P.niter)   // the next record: on its way to L1 while this iteration runs
      for (int o = lane * 16; o < S; o += 32 * 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + S + o));
    const double u_sel = rec[0], logu = rec[1], g1 = rec[2], g2 = rec[3];
    const uint2 wd = *reinterpret_cast<const uint2 *>(rec + 4);
    double A1[1][4] = {{0, 0, 0, 0}}, B1[1][4] = {{0, 0, 0, 0}}, A2[1][4] = {{0, 0, 0, 0}}, B2[1][4] = {{0, 0, 0, 0}};
    if (i0 < ld) {
      if (pt < k) {
        const double *r1 = rec + 8 + (size_t)pt * 2 * ld + i0;
        const double2 a = *reinterpret_cast<const double2 *>(r1), b = *reinterpret_cast<const double2 *>(r1 + 2);
        const double2 e = *reinterpret_cast<const double2 *>(r1 + ld), f = *reinterpret_cast<const double2 *>(r1 + ld + 2);
        A1[0][0] = a.x; A1[0][1] = a.y; A1[0][2] = b.x; A1[0][3] = b.y; B1[0][0] = e.x; B1[0][1] = e.y; B1[0][2] = f.x; B1[0][3] = f.y;
      }
      if (pt < k - 1) {
        const double *r2 = rec + 8 + (size_t)(k + pt) * 2 * ld + i0;
        const double2 a = *reinterpret_cast<const double2 *>(r2), b = *reinterpret_cast<const double2 *>(r2 + 2);
        const double2 e = *reinterpret_cast<const double2 *>(r2 + ld), f = *reinterpret_cast<const double2 *>(r2 + ld + 2);
        A2[0][0] = a.x; A2[0][1] = a.y; A2[0][2] = b.x; A2[0][3] = b.y; B2[0][0] = e.x; B2[0][1] = e.y; B2[0][2] = f.x; B2[0][3] = f.y;
      }
    }
    Decisions dc;
    dc.run_snooker = (int)(wd.x & 1u); dc.cr_idx = (int)((wd.x >> 1) & 15u); dc.lvl_idx = (int)((wd.x >> 5) & 15u);
    dc.delta = (int)((wd.x >> 9) & 15u);
    dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
    Stream s; s.init(P.cfg.seed, c_global, (uint32_t)iter);      // (only the boundary redraws draw here)
    MtOutcome o;
    double q[1][4] = {{x0[0][0], x0[0][1], x0[0][2], x0[0][3]}};
    bool fast = true;
#pragma unroll 1
    for (int bt = 0; bt < 2; ++bt) {   // the proposals around x0, then the reference set around the selected one (one copy of the code)
      if (bt) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { A1[0][j] = A2[0][j]; B1[0][j] = B2[0][j]; }
      }
      mt2_batch<G>(c, s, dc.run_snooker != 0, bt ? g2 : g1, bt ? k - 1 : k, pt, q, A1, B1, slot, bt ? rpri : pri, bt ? rlik : lik,
                   bt ? rsnk : snk);
      if (bt == 0) {
        bool anyfinite = false;
        for (int p = 0; p < k; ++p) anyfinite |= isfinite(Tc * lik[p] + pri[p]);
        if (!anyfinite) { fast = false; break; }
        o.sel = mt_choose(pri, lik, Tc, k, u_sel, lane);
        load_slot<G, 1>(c, c.slots + (size_t)o.sel * ld, q);
        o.new_prior = pri[o.sel]; o.new_like = lik[o.sel];
        __syncwarp();   // every lane-group has read the selected point before the reference set overwrites the slots
      }
    }
    if (fast) {
      o.gamma_one = dc.run_snooker ? g2 == 1.0 : ((wd.y >> k) & ((1u << (k - 1)) - 1u)) != 0u;
      const double mr = mt_ratio(pri, lik, snk, rpri, rlik, rsnk, last_prior, last_like, Tc, k, dc.run_snooker != 0, lane);
      o.accepted = isfinite(mr) && logu < mr;                                        // metrop_select, Dream.py:980-998
    } else {
      __syncwarp();
      mtp_iteration<G>(c, iter, c_global, pt, slot, x0, last_prior, last_like, Tc, M, dc, q, o);
    }
    mt_commit<G>(c, P, it, iter, c_local, c_global, writer, M, dc, o, q, x0, last_prior, last_like, Tc);
  }
  if (writer) {
    store_row<G, 1>(c, P.st.X + (size_t)c_local * ld, x0);
    if (c.g == 0) { P.st.last_prior[c_local] = last_prior; P.st.last_like[c_local] = last_like; }
  }
}

template <int G>
int launch_mtp(const StepParams &P, size_t smem, cudaStream_t stream) {
  const int threads = 128, chains_per_cta = threads / 32;
  const int grid = (P.cfg.nchains_local + chains_per_cta - 1) / chains_per_cta;
  const bool two_stage = P.st.draw_ws != nullptr;
  auto kern = two_stage ? dreamzs_mtchain_kernel<G> : dreamzs_mtp_kernel<G>;
  if (smem > 48 * 1024) {
    static size_t smem_set[2][64] = {{0}};
    if (ensure_dynamic_smem(kern, smem, smem_set[two_stage ? 1 : 0]) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  }
  if (two_stage) {
    const int64_t ncts = (int64_t)P.cfg.nchains_local * P.niter, units = ncts * (2 * P.cfg.multitry - 1);
    dreamzs_mtdraw_scalars_kernel<<<(unsigned)((ncts + 63) / 64), 256, 0, stream>>>(P);
    dreamzs_mtdraw_kernel<G><<<(unsigned)((units + 256 / G - 1) / (256 / G)), 256, 0, stream>>>(P);
    if (cudaGetLastError() != cudaSuccess) return DREAMZS_E_LAUNCH;
  }
  kern<<<grid, threads, smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
