// Two-stage single-try step for the generic targets (BASELINE config C4: 200-D banana, 8192 chains), sm_100a.
//
// As in the multi-try kernels (dreamzs_mtp_kernel.cuh), everything an iteration of Dream.astep draws is a function of the
// Philox counters and of the archive, not of the chain state: decisions, archive rows, e, zeta, the crossover mask and
// hence d' and gamma, the Metropolis uniform (pydream/Dream.py:542-726, 980-998).  The fused generic kernel
// (dreamzs_step_kernel.cuh) makes them inside each chain's serial loop: one lane-group per chain, ~12 dependent Philox
// blocks + a Box-Muller + two HBM gathers per iteration on the chain's critical path, and only N chains' worth of warps
// in flight.  Here
//   dreamzs_stdraw_kernel<G, R>   one lane-group per (chain, iteration) of a span of iterations -- nb times the warps, no
//                                 dependence between them -- leaves per pair a record in dreamzs_state.draw_ws:
//                                   [0] log of the Metropolis uniform  [1] snooker gamma  [2] two uint32: decisions (bit 0
//                                   snooker, 1-4 CR index, 5-8 gamma level, 9-12 DE pairs), gamma == 1 flag  [3] unused
//                                   [4 ..) A[ld], B[ld]: DE: J = (e*gamma)*diff and -- as float32 in the first half of B --
//                                   the normals of zeta, both 0 where the crossover keeps the centre; snooker: z and z1 - z2
//   dreamzs_stchain_kernel<G, R>  the Markov chains: proposal = (x + J) + zeta (or the snooker projection), bounds,
//                                 log prior + log-likelihood, Metropolis test against the precomputed log u, state / trace /
//                                 archive append -- no random numbers except the boundary redraws.
// The launcher walks a window in sub-spans when the scratch does not hold the records of a whole window (measured at C4: a
// whole window per kernel pair, 265 MB of records through HBM, beats L2-resident sub-spans of two iterations, 264 vs 209 M
// chain-steps/s: the launches cost more than the traffic).
// Draws, arithmetic and decisions are those of dreamzs_step_kernel<G, R, false> / the oracle (shared device functions).
#pragma once
#include "dreamzs_step_kernel.cuh"

namespace dreamzs {

__host__ __device__ inline int st2_record_doubles(int ld) { return 4 + 2 * ld; }

template <int G, int R>
__global__ void __launch_bounds__(256) dreamzs_stdraw_kernel(const __grid_constant__ StepParams P) {
  const int d = P.cfg.ndim, ld = P.cfg.ld, wn = P.niter;
  const int S = st2_record_doubles(ld);
  const int tid = threadIdx.x, lane = tid & 31;
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;
  if (P.wait_k) {   // sharded archive: the peers' rows of the previous append must have landed in this replica
    __shared__ int ok_s;
    if (tid == 0) ok_s = peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error) ? 1 : 0;
    __syncthreads();
    if (!ok_s) return;
  }
  Ctx<G, R> c{P, nullptr, nullptr, nullptr, G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1))), lane & (G - 1), d, ld};
  const int64_t unit = (int64_t)blockIdx.x * (256 / G) + tid / G;
  if (unit >= (int64_t)P.cfg.nchains_local * wn) return;     // (whole lane-groups leave together)
  const int c_local = (int)(unit / wn), itb = (int)(unit - (int64_t)c_local * wn);
  double *rec = P.st.draw_ws + (size_t)unit * S;
  Stream s; s.init(P.cfg.seed, (uint32_t)(P.cfg.chain_begin + c_local), (uint32_t)(P.iter_begin + itb));
  Decisions dc; dc.run_snooker = 0;
  const Bases b = {0u, 0u, 0u};
  double A[R][4], B[R][4], gamma = 0.0, logu;
  float NF[R][4];
  bool gone = false;
  if constexpr (G == 32) {
    // A warp per pair: the ten scalar draws of the iteration in ONE Philox pass, a draw per lane (a warp instruction costs
    // the same whether one lane needs it or 32): 0 snooker, 1 CR, 2 gamma level (multinomial calls 0, s0, s0+1), 3 DE pairs
    // (randint), 4 gamma unity (multinomial call s0+2), 5-6 np.random.uniform() calls 0-1, 7-9 random.sample calls 0-2
    const uint32_t s0 = P.cfg.snooker != 0 ? 1u : 0u;
    uint32_t call = 0, st = ST_MULTINOMIAL;
    if (lane == 1) call = s0;
    else if (lane == 2) call = s0 + 1;
    else if (lane == 3) st = ST_RANDINT;
    else if (lane == 4) call = s0 + 2;
    else if (lane == 5 || lane == 6) { st = ST_UNIFORM_SCAL; call = (uint32_t)(lane - 5); }
    else if (lane >= 7) { st = ST_SAMPLE; call = (uint32_t)(min(lane, 9) - 7); }
    const uint4 w = s.block(call, (int)st, 0);
    auto word = [&](int k) { return make_uint4(__shfl_sync(0xffffffffu, w.x, k), __shfl_sync(0xffffffffu, w.y, k),
                                               __shfl_sync(0xffffffffu, w.z, k), __shfl_sync(0xffffffffu, w.w, k)); };
    auto u53k = [&](int k) { return u53_of(__shfl_sync(0xffffffffu, w.x, k), __shfl_sync(0xffffffffu, w.y, k)); };
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
    if (s0) dc.run_snooker = u53k(0) < 0.0 + P.cfg.snooker;
    dc.cr_idx = invcdf(u53k(1), P.st.cr_probs, P.cfg.nCR);
    dc.lvl_idx = invcdf(u53k(2), P.st.gamma_probs, P.cfg.ngamma);
    dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
    dc.delta = 1;
    if (P.cfg.nDEpairs > 1) dc.delta = 1 + (int)(((uint64_t)__shfl_sync(0xffffffffu, w.x, 3) * (uint64_t)P.cfg.nDEpairs) >> 32);
    s.n_multinomial = s0 + 2; s.n_randint = P.cfg.nDEpairs > 1 ? 1u : 0u;
    const double u5 = u53k(5), u6 = u53k(6);
    if (dc.run_snooker) {
      gamma = 1.2 + (2.2 - 1.2) * u5;                        // Dream.py:618 (the unity draw before it is discarded)
      gone = gamma == 1.0;
      const uint4 pre[3] = {word(7), word(8), word(9)};
      snooker_rows<G, R>(c, s, b, 1, 0, P.archive_rows, A, B, pre);
      logu = log(u6);
    } else {
      unsigned reset;
      const uint4 ws = word(7);
      const double uu = u53k(4);
      de_draw<G, R>(c, s, dc, b, 1, 0, P.archive_rows, A, B, reset, gone, &ws, &uu, NF);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (((reset >> (4 * r + j)) & 1u) || c.dim0(r) + j >= d) { A[r][j] = 0.0; NF[r][j] = 0.0f; }
      logu = log(u5);
    }
  } else {
    if (P.cfg.snooker != 0) dc.run_snooker = multinomial2(s, P.cfg.snooker) == 0;      // set_snooker, Dream.py:542-554
    dc.cr_idx = multinomial_index(s, P.st.cr_probs, P.cfg.nCR);                        // set_CR, :556-569
    dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
    dc.delta = 1;
    if (P.cfg.nDEpairs > 1) {                                                          // set_DEpair, :571-583
      const uint4 w = s.block(s.n_randint++, ST_RANDINT, 0);
      dc.delta = 1 + (int)(((uint64_t)w.x * (uint64_t)P.cfg.nDEpairs) >> 32);
    }
    dc.lvl_idx = multinomial_index(s, P.st.gamma_probs, P.cfg.ngamma);                 // set_gamma_level, :585-599
    if (dc.run_snooker) {
      (void)multinomial2(s, P.cfg.p_gamma_unity);              // drawn and discarded (Dream.py:615-618)
      gamma = 1.2 + (2.2 - 1.2) * uniform_scalar(s);
      gone = gamma == 1.0;
      snooker_rows<G, R>(c, s, b, 1, 0, P.archive_rows, A, B);
    } else {
      unsigned reset;
      de_draw<G, R>(c, s, dc, b, 1, 0, P.archive_rows, A, B, reset, gone, nullptr, nullptr, NF);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (((reset >> (4 * r + j)) & 1u) || c.dim0(r) + j >= d) { A[r][j] = 0.0; NF[r][j] = 0.0f; }
    }
    logu = log(uniform_scalar(s));                // metrop_select's uniform: the next np.random.uniform() call
  }
  if (c.g == 0) {
    rec[0] = logu;
    rec[1] = gamma;
    const uint32_t dec = (dc.run_snooker ? 1u : 0u) | ((uint32_t)dc.cr_idx << 1) | ((uint32_t)dc.lvl_idx << 5) | ((uint32_t)dc.delta << 9);
    *reinterpret_cast<uint2 *>(rec + 2) = make_uint2(dec, gone ? 1u : 0u);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = c.dim0(r);
    if (i0 < ld) {
      double *rp = rec + 4 + i0;
      *reinterpret_cast<double2 *>(rp) = make_double2(A[r][0], A[r][1]);
      *reinterpret_cast<double2 *>(rp + 2) = make_double2(A[r][2], A[r][3]);
      if (dc.run_snooker) {
        *reinterpret_cast<double2 *>(rp + ld) = make_double2(B[r][0], B[r][1]);
        *reinterpret_cast<double2 *>(rp + ld + 2) = make_double2(B[r][2], B[r][3]);
      } else *reinterpret_cast<float4 *>(reinterpret_cast<float *>(rec + 4 + ld) + i0) = make_float4(NF[r][0], NF[r][1], NF[r][2], NF[r][3]);
    }
  }
}

#ifndef DZ_ST2_MINBLOCKS
#define DZ_ST2_MINBLOCKS 6
#endif
template <int G, int R>
__global__ void __launch_bounds__(128, (R <= 2 ? DZ_ST2_MINBLOCKS : 1)) dreamzs_stchain_kernel(const __grid_constant__ StepParams P) {
  extern __shared__ __align__(16) double smem[];
  const int d = P.cfg.ndim, ld = P.cfg.ld;
  const double *table = P.st.target_table;
  double *sm_chain = smem;
  if (P.table_in_smem) {
    for (int i = threadIdx.x; i < P.table_doubles; i += blockDim.x) smem[i] = table[i];
    table = smem;
    sm_chain = smem + ((P.table_doubles + 1) & ~1);
    __syncthreads();
  }
  constexpr int CHAINS_PER_WARP = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chain_in_cta = warp * CHAINS_PER_WARP + lane / G;
  const int chains_per_cta = (blockDim.x >> 5) * CHAINS_PER_WARP;
  const int c_local = blockIdx.x * chains_per_cta + chain_in_cta;
  if (c_local >= P.cfg.nchains_local) return;
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // the draw kernel's wait for the peers timed out
  Ctx<G, R> c{P, table, sm_chain + (size_t)chain_in_cta * ld, nullptr,
              G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1))), lane & (G - 1), d, ld};
  const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);
  double x0[R][4];
  load_slot<G, R>(c, P.st.X + (size_t)c_local * ld, x0);
  double last_prior = P.st.last_prior[c_local], last_like = P.st.last_like[c_local];
  const double Tc = P.temperature ? P.temperature[c_local] : 1.0;
  const int64_t M = P.archive_rows;
  const int S = st2_record_doubles(ld);
  const double *rec = P.st.draw_ws + (size_t)c_local * P.niter * S;
#pragma unroll 1
  for (int it = 0; it < P.niter; ++it, rec += S) {
    const int64_t iter = P.iter_begin + it;
    if (it + 1 < P.niter)   // the next record: on its way to L1 while this iteration runs
      for (int o = c.g * 16; o < S; o += G * 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + S + o));
    const double logu = rec[0], gamma = rec[1];
    const uint2 wd = *reinterpret_cast<const uint2 *>(rec + 2);
    const int run_snooker = (int)(wd.x & 1u);
    double A[R][4], B[R][4], q[R][4];
    load_slot<G, R>(c, rec + 4, A);
    double snk0 = 0.0, D0 = 0.0;
    if (run_snooker) {
      load_slot<G, R>(c, rec + 4 + ld, B);
      snooker_compute<G, R>(c, 1, gamma, x0, A, B, q, snk0, D0);
    } else {
      const float *nf = reinterpret_cast<const float *>(rec + 4 + ld);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i0 = c.dim0(r);
        const float4 n4 = i0 < ld ? *reinterpret_cast<const float4 *>(nf + i0) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)   // Dream.py:694, 717: zeta = np.random.normal(0, zeta); the centre is kept where J = zeta = 0
          q[r][j] = x0[r][j] + A[r][j] + (0.0 + P.cfg.zeta * (double)nn[j]);
      }
    }
    if (P.cfg.hardboundaries && !P.all_flat) {
      Stream s; s.init(P.cfg.seed, c_global, (uint32_t)iter);      // (only the boundary redraws draw here)
      apply_bounds<G, R>(c, s, q);
    }
    __syncwarp(c.gmask);
    store_slot<G, R>(c, c.slots, q);
    __syncwarp(c.gmask);
    double q_prior, q_like;
    eval_logp<G, R>(c, q, c.slots, q_prior, q_like);
    const double last_logp = Tc * last_like + last_prior, q_logp = Tc * q_like + q_prior;
    double mr;
    if (run_snooker) {                                                               // Dream.py:326-332
      const double norm = sqrt(D0);
      const double cur = (norm != 0 ? log(norm) : 0.0) * (d - 1);
      mr = nan_to_num((q_logp + snk0) - (last_logp + cur));
    } else mr = nan_to_num(q_logp) - nan_to_num(last_logp);                          // Dream.py:334
    const bool accepted = isfinite(mr) && logu < mr;                                 // metrop_select, :980-998
    int changed = 0;
    if (accepted) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) { changed |= (q[r][j] != x0[r][j]); x0[r][j] = q[r][j]; }
    }
    changed = gsum_int<G>(changed, c.gmask) != 0;
    if (changed) { last_prior = q_prior; last_like = q_like; }
    const int64_t trow = P.tr.trace_offset + it;
    store_row<G, R>(c, P.tr.trace + ((size_t)c_local * P.tr.trace_iters + trow) * ld, x0);
    if (c.g == 0) {
      P.tr.trace_logp[(size_t)c_local * P.tr.trace_iters + trow] = Tc * last_like + last_prior;   // core.py:115 (T = 1), :176
      if (P.tr.decisions)
        P.tr.decisions[(size_t)c_local * P.tr.trace_iters + trow] =
            pack_decision(changed, run_snooker, (int)((wd.x >> 1) & 15u), (int)((wd.x >> 5) & 15u), (int)((wd.x >> 9) & 15u), 0,
                          wd.y != 0u, accepted);
    }
    if (iter % P.cfg.history_thin == 0) {   // record_history: only the last iteration of a window appends
      store_row<G, R>(c, P.st.Z + (size_t)(M + c_global) * ld, x0);
      for (int pz = 0; pz < P.npeers; ++pz) store_row<G, R>(c, P.peer_Z[pz] + (size_t)(M + c_global) * ld, x0);   // replicas over NVLink
      if (P.publish_k) {
        __threadfence_system();
        __syncwarp(c.gmask);
        if (c.g == 0) peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
      }
    }
    __syncwarp(c.gmask);
  }
  store_row<G, R>(c, P.st.X + (size_t)c_local * ld, x0);
  if (c.g == 0) { P.st.last_prior[c_local] = last_prior; P.st.last_like[c_local] = last_like; }
}

// A window in sub-spans whose records fit the scratch (sized by the caller so that they stay in L2): draw kernel, chain
// kernel, next sub-span.  Only the window's last iteration may append, so only the last sub-span publishes to the peers;
// only the first waits for them.
template <int G, int R>
int launch_st2(const StepParams &P0, int threads, size_t smem, cudaStream_t stream) {
  const int S = st2_record_doubles(P0.cfg.ld);
  const int64_t per_iter = (int64_t)P0.cfg.nchains_local * S * (int64_t)sizeof(double);
  int nb = (int)(P0.st.draw_ws_bytes / per_iter);
  if (nb < 1) return DREAMZS_E_BADARG;
  auto chain = dreamzs_stchain_kernel<G, R>;
  if (smem > 48 * 1024) {
    static size_t smem_set[64] = {0};
    if (ensure_dynamic_smem(chain, smem, smem_set) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  }
  const int chains_per_cta = (threads / 32) * (32 / G);
  const int grid = (P0.cfg.nchains_local + chains_per_cta - 1) / chains_per_cta;
  for (int t0 = 0; t0 < P0.niter; t0 += nb) {
    StepParams P = P0;
    P.iter_begin = P0.iter_begin + t0;
    P.niter = P0.niter - t0 < nb ? P0.niter - t0 : nb;
    P.tr.trace_offset = P0.tr.trace_offset + t0;
    if (t0 > 0) P.wait_k = 0;
    if (t0 + P.niter < P0.niter) P.publish_k = 0;
    const int64_t units = (int64_t)P.cfg.nchains_local * P.niter;
    dreamzs_stdraw_kernel<G, R><<<(unsigned)((units + 256 / G - 1) / (256 / G)), 256, 0, stream>>>(P);
    chain<<<grid, threads, smem, stream>>>(P);
    if (cudaGetLastError() != cudaSuccess) return DREAMZS_E_LAUNCH;
  }
  return DREAMZS_OK;
}

}  // namespace dreamzs
