// Window kernel for the dense-Gaussian target (BASELINE config C2), sm_100a.
//
// The CTA-synchronous kernel (dreamzs_gauss_kernel.cuh) keeps the whole of Dream.astep on the
// critical path of every iteration: ~17 k cycles per iteration of which ~1.1 k are the quadratic
// forms.  This kernel takes everything that does not depend on the chain state off that path.
//
// For a flat prior and one DE pair the jump of iteration t,  dx = (e*gamma)*(z_r1 - z_r2) + zeta
// with its crossover mask, the snooker rows, the accept uniform -- everything except the current
// state x -- is a function of the Philox counters and of the archive, which is constant inside a
// launch.  And with y = invC x and Q = x.y carried per chain, the quadratic form of a proposal is
//     Q(x + dx) = Q + 2 dx.y + dx.(invC dx),
// where w = invC dx does not depend on x either.  So a launch works in batches of NB iterations;
// a (chain, iteration) pair is a "column":
//   G  (warp per column)  all scalar draws of a warp's columns in ONE lane-parallel Philox pass;
//                         archive rows TMA-staged straight into the shared-memory slots that will
//                         hold the jump (UBLKCP + mbarrier complete_tx); dx, zeta, crossover mask
//   M  (warp per tile)    W = invC * [dx columns | snooker z columns | x column on refresh]:
//                         register-tiled fp64 products, tile = one chain's columns x 4 rows per
//                         lane, K split 2-fold over warps
//   C  (warp per chain)   the Markov chain itself: per iteration 2 adds, one 4-wide dot product and
//                         ONE warp reduction, the Metropolis test, trace write, archive append
// The three phases stress different units (G: integer pipe + fp64 transcendentals, M: fp64 FMA +
// shared memory, C: latency), so the CTA is split into GW_GROUPS independent warp groups, each
// owning a share of the CTA's chains and running G -> M -> C on its own named barrier; the groups
// drift apart and one group's M overlaps the other's G or C.  The precision matrix (80 KB at
// d=100) is shared by the groups.
// The snooker move is linear in the state as well: dx = c (x - z), invC dx = c (y - invC z), so
// its column of M is invC z.  y and Q are refreshed from x every DREAMZS_GAUSS_REFRESH_WINDOWS
// windows so rounding drift stays orders of magnitude below the 1e-12 parity tolerance.
//
// RNG consumption, decisions and element-wise arithmetic are those of dreamzs_step_kernel / the
// oracle; only the summation order of the quadratic form differs.
#pragma once
#include "dreamzs_gauss_kernel.cuh"

namespace dreamzs {

constexpr int GW_THREADS = 512;
#ifndef DZ_GW_GROUPS
#define DZ_GW_GROUPS 2
#endif
constexpr int GW_GROUPS = DZ_GW_GROUPS;                    // independent warp groups per CTA (2 or 4; A/B builds override)
// A/B builds only (default 0 = off): the odd groups run a short first batch of DZ_GW_STAGGER iterations, so that from
// then on the groups of a CTA are out of phase (one group's latency-bound C next to the other's issue-bound G / M;
// DESIGN.md section 9).  Results do not depend on how a launch is cut into batches.
#ifndef DZ_GW_STAGGER
#define DZ_GW_STAGGER 0
#endif
constexpr int GW_GWARPS = GW_THREADS / 32 / GW_GROUPS;     // warps per group
constexpr int GW_GTHREADS = GW_GWARPS * 32;
constexpr int GW_KS = 2;       // K split of the products; fixed: the summation order is part of the result
constexpr int GW_MAXNB = 5;    // iterations per batch
constexpr int GW_MAXCOLW = 3;  // columns a warp generates per batch (3 x 10 lanes of scalar draws)

__host__ __device__ inline int gwin_group_chains(int TC, int g) { return (TC + GW_GROUPS - 1 - g) / GW_GROUPS; }
__host__ __device__ inline int gwin_group_first(int TC, int g) {
  int b = 0;
  for (int i = 0; i < g; ++i) b += gwin_group_chains(TC, i);
  return b;
}

struct GwinLayout {
  int d2, ncol, gcap;
  size_t oAt, oWc, oJc, oZc, oXs, oYs, oGam, oLogu, oGsn, oProbs, oMbar, oMeta, bytes;
};

__host__ __device__ inline GwinLayout gwin_layout(int d, int ld, int TC, int NB, int ngamma) {
  GwinLayout L;
  L.d2 = (d + 1) & ~1;
  L.gcap = gwin_group_chains(TC, 0) * NB;            // column slots per group (chain-major: chain * NB + iteration)
  L.ncol = L.gcap * GW_GROUPS;
  size_t o = 0;
  L.oAt = o;    o += (size_t)L.d2 * ld;              // precision matrix, transposed, row stride ld
  L.oWc = o;    o += (size_t)(L.ncol + TC) * ld;     // dx / z columns -> invC * column; + TC refresh columns
  L.oJc = o;    o += (size_t)L.ncol * ld;            // (e*gamma)*diff      | snooker: z
  L.oZc = o;    o += (size_t)L.ncol * ld;            // zeta                | snooker: z1 - z2
  L.oXs = o;    o += (size_t)TC * ld;                // chain states between the C phases
  L.oYs = o;    o += (size_t)TC * ld;
  L.oGam = o;   o += ((size_t)ngamma * d + 1) & ~(size_t)1;
  L.oLogu = o;  o += L.ncol;
  L.oGsn = o;   o += L.ncol;
  L.oProbs = o; o += 32 + 4 * TC;
  L.oMbar = o;  o += L.ncol + 1;
  L.oMeta = o;  o += (L.ncol + 1) / 2;
  L.bytes = o * sizeof(double);
  return L;
}

__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void group_sync(int gid) {
  asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(GW_GTHREADS) : "memory");
}

template <int TC>
__global__ void __launch_bounds__(GW_THREADS, 1) dreamzs_gwin_kernel(const StepParams P) {
  extern __shared__ __align__(16) double smem[];
  const int d = P.cfg.ndim, ld = P.cfg.ld, NB = P.gw_nb;
  const GwinLayout L = gwin_layout(d, ld, TC, NB, P.cfg.ngamma);
  const int d2 = L.d2, ncol = L.ncol;
  double *At = smem + L.oAt, *Wc = smem + L.oWc, *Jc = smem + L.oJc, *Zc = smem + L.oZc;
  double *Xs = smem + L.oXs, *Ys = smem + L.oYs, *gam = smem + L.oGam;
  double *logu = smem + L.oLogu, *gsn = smem + L.oGsn;
  double *probs = smem + L.oProbs, *cst = probs + 32;   // cst: per chain [Q, last_prior, last_like, -]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem + L.oMbar);   // [0, ncol) columns, [ncol] precision matrix
  uint32_t *meta = reinterpret_cast<uint32_t *>(smem + L.oMeta);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = warp / GW_GWARPS, gw = warp - gid * GW_GWARPS;    // warp group, warp within the group
  const double logF = P.st.target_table[0];
  int dbg_n = 0;
#if DZ_GW_STAGGER > 0   // A/B builds: also stamp warp 0 of the second group (entries 32..63), to see the phase offset
#define GW_STAMP() do { if (P.dbg && blockIdx.x == 0 && (tid == 0 || tid == GW_GTHREADS)) P.dbg[(tid ? 32 : 0) + dbg_n] = clock64(); ++dbg_n; } while (0)
#else
#define GW_STAMP() do { if (P.dbg && blockIdx.x == 0 && tid == 0) P.dbg[dbg_n] = clock64(); ++dbg_n; } while (0)
#endif
  GW_STAMP();   // 0: kernel entry

  const int i0 = 4 * lane;
  const bool own = i0 < ld;          // lane owns a 4-dimension chunk of a row
  const int64_t M = P.archive_rows;
  const uint32_t row_bytes = (uint32_t)ld * 8u;
  const uint32_t s0 = P.cfg.snooker != 0 ? 1u : 0u;   // multinomial call number of the CR draw
  const uint32_t k0 = (uint32_t)P.cfg.seed, k1 = (uint32_t)(P.cfg.seed >> 32);
  const int cta_chain0 = blockIdx.x * TC;
  const int nch_cta = min(TC, P.cfg.nchains_local - cta_chain0);   // chains of this CTA
  const int gfirst = gwin_group_first(TC, gid);                     // first chain slot of this group
  const int gch = max(0, min(gwin_group_chains(TC, gid), nch_cta - gfirst));   // chains of this group
  const int gbase = gid * L.gcap;                                   // first column slot of this group

  // ---- prologue: one TMA bulk copy brings the precision matrix; chain states -> shared memory
  if (tid < ncol + 1) mbar_init(mbar + tid, 1);
  if (tid < 32) {
    double v = 0.0;
    if (tid < 16) v = tid < P.cfg.nCR ? P.st.cr_probs[tid] : 0.0;
    else if (tid < 24) v = tid - 16 < P.cfg.ngamma ? P.st.gamma_probs[tid - 16] : 0.0;
    else if (tid == 24) v = P.cfg.snooker;
    else if (tid == 26) v = P.cfg.p_gamma_unity;
    probs[tid] = v;
  }
  for (int i = tid; i < P.cfg.ngamma * d; i += GW_THREADS) {   // gamma_table[level][0][:] (one DE pair)
    const int lv = i / d;
    gam[i] = P.st.gamma_table[(size_t)lv * P.cfg.nDEpairs * d + (i - lv * d)];
  }
  if ((d & 1) && tid < ld) At[(size_t)d * ld + tid] = 0.0;     // padding row of the j-pair loop
  if (warp < nch_cta) {
    const int c_local = cta_chain0 + warp;
    if (own) {
      const double *xrow = P.st.X + (size_t)c_local * ld + i0;
      *reinterpret_cast<double2 *>(Xs + warp * ld + i0) = *reinterpret_cast<const double2 *>(xrow);
      *reinterpret_cast<double2 *>(Xs + warp * ld + i0 + 2) = *reinterpret_cast<const double2 *>(xrow + 2);
      if (!P.gw_refresh) {
        const double *yrow = P.st.gauss_Y + (size_t)c_local * ld + i0;
        *reinterpret_cast<double2 *>(Ys + warp * ld + i0) = *reinterpret_cast<const double2 *>(yrow);
        *reinterpret_cast<double2 *>(Ys + warp * ld + i0 + 2) = *reinterpret_cast<const double2 *>(yrow + 2);
      }
    }
    if (lane == 0) {
      cst[warp * 4 + 0] = P.gw_refresh ? 0.0 : P.st.gauss_Q[c_local];
      cst[warp * 4 + 1] = P.st.last_prior[c_local];
      cst[warp * 4 + 2] = P.st.last_like[c_local];
    }
  }
  if (P.wait_k && tid == 32) peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error);   // peers' rows have landed
  __syncthreads();
  if (tid == 0) {
    fence_proxy_async();
    const uint32_t bytes = (uint32_t)d * row_bytes;
    mbar_expect_tx(mbar + ncol, bytes);
    tma_load_row(At, P.st.target_table + 2, bytes, mbar + ncol);
  }
  GW_STAMP();   // 1: prologue done

  const int nq = ld / 4;                                   // row groups of a tile = lanes of a tile warp
  const int jl = ((d2 / 2 + GW_KS - 1) / GW_KS) * 2;       // K range of a split (even)

  int done = 0;
#if DZ_GW_STAGGER > 0
  uint32_t colphase = 0;   // bit k: mbarrier phase of this warp's k-th column (a column skipped by a short batch keeps its phase)
#endif
  for (int batch = 0; done < P.niter; ++batch) {
#if DZ_GW_STAGGER > 0
    int nb = min(NB, P.niter - done);
    if ((gid & 1) && batch == 0 && P.niter > DZ_GW_STAGGER) nb = min(nb, DZ_GW_STAGGER);
#else
    const int nb = min(NB, P.niter - done);
#endif
    const bool do_refresh = P.gw_refresh && batch == 0;
#if DZ_GW_STAGGER == 0
    const uint32_t parity = (uint32_t)(batch & 1);
#endif
    // ================================================================ G: generation (warp per column)
    // group-local column index cl = chain * NB + iteration; the warp owns cl = gw, gw + 8, gw + 16
    {
      // ---- S: scalar draws and archive rows of the warp's columns in one pass, 10 lanes per column:
      //      0-5: snooker, CR, gamma level, gamma unity (Dream.py:542-599, 615), first two np.random.uniform()
      //      (snooker gamma :618 / Metropolis :993); 6-8: random.sample calls 0-2 (sample_from_history, :646-668)
      const int sk = min(lane / 10, GW_MAXCOLW - 1), kind = lane - 10 * (lane / 10);
      {
        const int cl = gw + GW_GWARPS * sk;
        const int ch = cl / NB, itb = cl - ch * NB;
        const bool ok = lane < 10 * GW_MAXCOLW && ch < gch && itb < nb;
        const uint32_t iter = (uint32_t)(P.iter_begin + done + (ok ? itb : 0));
        const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + cta_chain0 + gfirst + (ok ? ch : 0));
        uint32_t call = 0, st = ST_MULTINOMIAL;
        const double *pp = probs + 24;
        int n = 2;
        if (kind == 1) { call = s0; pp = probs; n = P.cfg.nCR; }
        else if (kind == 2) { call = s0 + 1; pp = probs + 16; n = P.cfg.ngamma; }
        else if (kind == 3) { call = s0 + 2; pp = probs + 26; }
        else if (kind == 4) { st = ST_UNIFORM_SCAL; }
        else if (kind == 5) { call = 1; st = ST_UNIFORM_SCAL; }
        else if (kind >= 6) { call = (uint32_t)(kind - 6); st = ST_SAMPLE; }
        const uint4 w = philox4x32(0u, (call << 3) | st, iter, c_global, k0, k1);
        const double u = u53_of(w.x, w.y);
        double acc = 0.0;
        int idx = n - 1;
        bool found = false;
        for (int j = 0; j < n; ++j) {
          acc = acc + pp[j];
          if (!found && u < acc) { idx = j; found = true; }
        }
        const double lg = log(u);
        int64_t r0 = (int64_t)(((uint64_t)w.x * (uint64_t)M) >> 32);
        int64_t r1 = (int64_t)(((uint64_t)w.y * (uint64_t)(M - 1)) >> 32);
        if (r1 >= r0) r1 += 1;
        // column decisions -> every lane of the column's 10
        const int base = 10 * sk;
        const int snk = (s0 != 0u) && __shfl_sync(0xffffffffu, idx, base) == 0;
        const int cr_s = __shfl_sync(0xffffffffu, idx, base + 1), lvl_s = __shfl_sync(0xffffffffu, idx, base + 2);
        const int unity_s = __shfl_sync(0xffffffffu, idx, base + 3);
        const int col = gbase + cl;
        if (ok) {
          // meta word: bits 0-3 CR index, 4-7 gamma level, 8 snooker, 9 gamma == 1 (set in V), 10 "not unity"
          if (kind == 0) meta[col] = (uint32_t)cr_s | ((uint32_t)lvl_s << 4) | (snk ? 256u : 0u) | (unity_s != 0 ? 1024u : 0u);
          if (kind == (snk ? 5 : 4)) logu[col] = lg;                               // Metropolis uniform: 2nd draw after a snooker gamma
          if (kind == 4) gsn[col] = 1.2 + (2.2 - 1.2) * u;                         // snooker gamma, Dream.py:618
          if (kind == 6) {
            fence_proxy_async();   // earlier generic-proxy accesses of the slots are ordered before the async writes
            mbar_expect_tx(mbar + col, row_bytes * (snk ? 3u : 2u));
          }
        }
        __syncwarp();
        if (ok) {
          // DE: z_r1 -> J slot, z_r2 -> zeta slot;   snooker: z -> J slot, z1 -> W slot, z2 -> zeta slot
          double *js = Jc + (size_t)col * ld, *zs = Zc + (size_t)col * ld, *ws = Wc + (size_t)col * ld;
          if (!snk) {
            if (kind == 6) {
              tma_load_row(js, P.st.Z + (size_t)r0 * ld, row_bytes, mbar + col);
              tma_load_row(zs, P.st.Z + (size_t)r1 * ld, row_bytes, mbar + col);
            }
          } else if (kind >= 6 && kind <= 8) {
            fence_proxy_async();
            tma_load_row(kind == 6 ? js : kind == 7 ? ws : zs, P.st.Z + (size_t)r0 * ld, row_bytes, mbar + col);
          }
        }
        __syncwarp();
      }
      GW_STAMP();   // +0: rows requested (warp 0)
      // ---- V: the columns (generate_proposal_points DE branch, Dream.py:688-726; snooker rows, :808-810)
#pragma unroll 1
      for (int k = 0; k < GW_MAXCOLW; ++k) {
        const int cl = gw + GW_GWARPS * k;
        const int ch = cl / NB, itb = cl - ch * NB;
        if (ch >= gch || itb >= nb) continue;
        const int col = gbase + cl;
        const uint32_t iter = (uint32_t)(P.iter_begin + done + itb);
        const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + cta_chain0 + gfirst + ch);
        const uint32_t mt = meta[col];
#if DZ_GW_STAGGER > 0
        const uint32_t parity = (colphase >> k) & 1u;
        colphase ^= 1u << k;
#endif
        const int snk = (mt >> 8) & 1, cr_idx = mt & 15, lvl_idx = (mt >> 4) & 15;
        double *js = Jc + (size_t)col * ld + i0, *zs = Zc + (size_t)col * ld + i0, *ws = Wc + (size_t)col * ld + i0;
        if (!snk) {
          double zeta[4] = {0, 0, 0, 0}, e[4] = {1, 1, 1, 1};
          unsigned reset = 15u;
          int dprime = 0;
          if (own && i0 < d) {
            double nz[4];
            normal4(philox4x32((uint32_t)lane, (0u << 3) | ST_NORMAL, iter, c_global, k0, k1), nz);
            const uint4 we = philox4x32((uint32_t)lane, (0u << 3) | ST_UNIFORM_VEC, iter, c_global, k0, k1);
            const uint4 wu = philox4x32((uint32_t)lane, (1u << 3) | ST_UNIFORM_VEC, iter, c_global, k0, k1);
            const uint32_t wev[4] = {we.x, we.y, we.z, we.w}, wuv[4] = {wu.x, wu.y, wu.z, wu.w};
            // U = w 2^-32 exactly, so U < CR <=> w < ceil(CR 2^32) and U > CR <=> w > floor(CR 2^32)
            const double CRs = ((double)(cr_idx + 1) / (double)P.cfg.nCR) * 4294967296.0;
            const uint64_t t_lt = (uint64_t)ceil(CRs), t_gt = (uint64_t)floor(CRs);
            reset = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              zeta[j] = 0.0 + P.cfg.zeta * nz[j];
              e[j] = (-P.cfg.lamb + (P.cfg.lamb - (-P.cfg.lamb)) * u32_of(wev[j])) + 1;
              if (i0 + j < d) {
                dprime += ((uint64_t)wuv[j] < t_lt);
                if ((uint64_t)wuv[j] > t_gt) reset |= 1u << j;
              } else reset |= 1u << j;
            }
          }
          dprime = __reduce_add_sync(0xffffffffu, dprime);
          double gamma = 1.0;
          if (mt & 1024u) gamma = gam[lvl_idx * d + (dprime >= 1 ? dprime - 1 : d - 1)];
          if (lane == 0 && gamma == 1.0) meta[col] = mt | 512u;
          mbar_wait(mbar + col, parity);
          if (own) {
            const double2 a01 = *reinterpret_cast<const double2 *>(js), a23 = *reinterpret_cast<const double2 *>(js + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(zs), b23 = *reinterpret_cast<const double2 *>(zs + 2);
            const double diff[4] = {a01.x - b01.x, a01.y - b01.y, a23.x - b23.x, a23.y - b23.y};
            double J[4], zt[4], dl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool keep = !((reset >> j) & 1u);
              J[j] = keep ? (e[j] * gamma) * diff[j] : 0.0;
              zt[j] = keep ? zeta[j] : 0.0;
              dl[j] = J[j] + zt[j];
            }
            *reinterpret_cast<double2 *>(js) = make_double2(J[0], J[1]); *reinterpret_cast<double2 *>(js + 2) = make_double2(J[2], J[3]);
            *reinterpret_cast<double2 *>(zs) = make_double2(zt[0], zt[1]); *reinterpret_cast<double2 *>(zs + 2) = make_double2(zt[2], zt[3]);
            *reinterpret_cast<double2 *>(ws) = make_double2(dl[0], dl[1]); *reinterpret_cast<double2 *>(ws + 2) = make_double2(dl[2], dl[3]);
          }
        } else {
          if (lane == 0 && gsn[col] == 1.0) meta[col] = mt | 512u;
          mbar_wait(mbar + col, parity);
          if (own) {
            const double2 z01 = *reinterpret_cast<const double2 *>(js), z23 = *reinterpret_cast<const double2 *>(js + 2);
            const double2 a01 = *reinterpret_cast<const double2 *>(ws), a23 = *reinterpret_cast<const double2 *>(ws + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(zs), b23 = *reinterpret_cast<const double2 *>(zs + 2);
            *reinterpret_cast<double2 *>(zs) = make_double2(a01.x - b01.x, a01.y - b01.y);
            *reinterpret_cast<double2 *>(zs + 2) = make_double2(a23.x - b23.x, a23.y - b23.y);
            *reinterpret_cast<double2 *>(ws) = z01; *reinterpret_cast<double2 *>(ws + 2) = z23;
          }
        }
      }
      if (do_refresh && gw < gch && own) {   // refresh column of chain gw: x
        double *ws = Wc + (size_t)(ncol + gfirst + gw) * ld + i0;
        *reinterpret_cast<double2 *>(ws) = *reinterpret_cast<const double2 *>(Xs + (gfirst + gw) * ld + i0);
        *reinterpret_cast<double2 *>(ws + 2) = *reinterpret_cast<const double2 *>(Xs + (gfirst + gw) * ld + i0 + 2);
      }
    }
    GW_STAMP();   // +1: columns of warp 0 generated
    if (batch == 0) mbar_wait(mbar + ncol, 0);   // precision matrix has landed
    group_sync(gid);
    GW_STAMP();   // +2: the group's columns generated
    // ================================================================ M: W = invC * columns (warp per tile)
    // tile = (chain of the group: its NB columns + its refresh column, K range); lane = row group:
    // rows 2l, 2l+1, ld/2+2l, ld/2+2l+1.  Columns that do not exist in this batch (short last batch, no
    // refresh) alias an existing one and are not written back.
    {
      const int tch = gw % gwin_group_chains(TC, 0), ks = gw / gwin_group_chains(TC, 0);
      const bool active = tch < gch && ks < GW_KS && lane < nq;
      const int ncc = nb + (do_refresh ? 1 : 0);
      double acc[4][GW_MAXNB + 1];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc <= GW_MAXNB; ++cc) acc[r][cc] = 0.0;
      uint32_t xaddr[GW_MAXNB + 1];       // shared-window byte addresses of the tile's columns
#pragma unroll
      for (int cc = 0; cc <= GW_MAXNB; ++cc) {
        int c = gbase + tch * NB + min(cc, nb - 1);                      // iteration cc of the chain
        if (cc >= nb) c = do_refresh ? ncol + gfirst + tch : gbase + tch * NB;   // its refresh column / an alias
        xaddr[cc] = smem_u32(Wc + (size_t)c * ld);
      }
      if (active) {
        const int jb = min(d2, ks * jl), je = min(d2, jb + jl);
        const uint32_t hb = (uint32_t)ld * 4u;                           // ld/2 doubles in bytes
        const uint32_t rowb = (uint32_t)ld * 8u;
        uint32_t ap = smem_u32(At + 2 * lane) + (uint32_t)jb * rowb;
        uint32_t joff = (uint32_t)jb * 8u;
        if (ncc == GW_MAXNB) {
#pragma unroll 1
          for (int j = jb; j < je; j += 2, ap += 2 * rowb, joff += 16) {
            const double2 a0 = lds_f64x2(ap), b0 = lds_f64x2(ap + hb), a1 = lds_f64x2(ap + rowb), b1 = lds_f64x2(ap + rowb + hb);
            double2 x[GW_MAXNB];
#pragma unroll
            for (int cc = 0; cc < GW_MAXNB; ++cc) x[cc] = lds_f64x2(xaddr[cc] + joff);
#pragma unroll
            for (int cc = 0; cc < GW_MAXNB; ++cc) {
              acc[0][cc] = fma(a0.x, x[cc].x, acc[0][cc]); acc[1][cc] = fma(a0.y, x[cc].x, acc[1][cc]);
              acc[2][cc] = fma(b0.x, x[cc].x, acc[2][cc]); acc[3][cc] = fma(b0.y, x[cc].x, acc[3][cc]);
            }
#pragma unroll
            for (int cc = 0; cc < GW_MAXNB; ++cc) {
              acc[0][cc] = fma(a1.x, x[cc].y, acc[0][cc]); acc[1][cc] = fma(a1.y, x[cc].y, acc[1][cc]);
              acc[2][cc] = fma(b1.x, x[cc].y, acc[2][cc]); acc[3][cc] = fma(b1.y, x[cc].y, acc[3][cc]);
            }
          }
        } else {
#pragma unroll 1
          for (int j = jb; j < je; j += 2, ap += 2 * rowb, joff += 16) {
            const double2 a0 = lds_f64x2(ap), b0 = lds_f64x2(ap + hb), a1 = lds_f64x2(ap + rowb), b1 = lds_f64x2(ap + rowb + hb);
            double2 x[GW_MAXNB + 1];
#pragma unroll
            for (int cc = 0; cc <= GW_MAXNB; ++cc) x[cc] = lds_f64x2(xaddr[cc] + joff);
#pragma unroll
            for (int cc = 0; cc <= GW_MAXNB; ++cc) {
              acc[0][cc] = fma(a0.x, x[cc].x, acc[0][cc]); acc[1][cc] = fma(a0.y, x[cc].x, acc[1][cc]);
              acc[2][cc] = fma(b0.x, x[cc].x, acc[2][cc]); acc[3][cc] = fma(b0.y, x[cc].x, acc[3][cc]);
            }
#pragma unroll
            for (int cc = 0; cc <= GW_MAXNB; ++cc) {
              acc[0][cc] = fma(a1.x, x[cc].y, acc[0][cc]); acc[1][cc] = fma(a1.y, x[cc].y, acc[1][cc]);
              acc[2][cc] = fma(b1.x, x[cc].y, acc[2][cc]); acc[3][cc] = fma(b1.y, x[cc].y, acc[3][cc]);
            }
          }
        }
      }
      group_sync(gid);   // every column of the group has been read: the products may now overwrite them in place
      GW_STAMP();   // +3: products computed
#pragma unroll 1
      for (int r = 0; r < GW_KS; ++r) {
        if (active && ks == r) {
          const uint32_t hb = (uint32_t)ld * 4u;
#pragma unroll
          for (int cc = 0; cc <= GW_MAXNB; ++cc) {
            if (cc < nb || (cc == nb && do_refresh)) {
              const uint32_t wp = xaddr[cc] + 16u * (uint32_t)lane;
              if (r > 0) {
                const double2 ta = lds_f64x2(wp), tb = lds_f64x2(wp + hb);
                acc[0][cc] = ta.x + acc[0][cc]; acc[1][cc] = ta.y + acc[1][cc];
                acc[2][cc] = tb.x + acc[2][cc]; acc[3][cc] = tb.y + acc[3][cc];
              }
              sts_f64x2(wp, acc[0][cc], acc[1][cc]);
              sts_f64x2(wp + hb, acc[2][cc], acc[3][cc]);
            }
          }
        }
        group_sync(gid);
      }
    }
    GW_STAMP();   // +4: products written
    // ================================================================ C: the chains (warp per chain)
    if (gw < gch) {
      const int cs = gfirst + gw;                 // chain slot in the CTA
      const int c_local = cta_chain0 + cs;
      const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);
      double x0[4] = {0, 0, 0, 0}, y0[4] = {0, 0, 0, 0};
      double Q0 = cst[cs * 4], last_prior = cst[cs * 4 + 1], last_like = cst[cs * 4 + 2];
      if (own) {
        const double2 a = *reinterpret_cast<const double2 *>(Xs + cs * ld + i0), b = *reinterpret_cast<const double2 *>(Xs + cs * ld + i0 + 2);
        x0[0] = a.x; x0[1] = a.y; x0[2] = b.x; x0[3] = b.y;
      }
      if (do_refresh) {
        const double *ws = Wc + (size_t)(ncol + cs) * ld + i0;
        double part = 0.0;
        if (own) {
          const double2 a = *reinterpret_cast<const double2 *>(ws), b = *reinterpret_cast<const double2 *>(ws + 2);
          y0[0] = a.x; y0[1] = a.y; y0[2] = b.x; y0[3] = b.y;
#pragma unroll
          for (int j = 0; j < 4; ++j) part = fma(x0[j], y0[j], part);
        }
        Q0 = gsum<32>(part, 0xffffffffu);
      } else if (own) {
        const double2 a = *reinterpret_cast<const double2 *>(Ys + cs * ld + i0), b = *reinterpret_cast<const double2 *>(Ys + cs * ld + i0 + 2);
        y0[0] = a.x; y0[1] = a.y; y0[2] = b.x; y0[3] = b.y;
      }
      double ntn_last = nan_to_num(1.0 * last_like + last_prior);
      // column data of the iteration about to run (loaded one iteration ahead)
      double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, w[4] = {0, 0, 0, 0};
      uint32_t mt = 0;
      double lu = 0.0, gsnk = 0.0, sd = 0.0;
      auto load_col = [&](int itb, double (&aa)[4], double (&bb)[4], double (&ww)[4], uint32_t &m, double &l, double &g) {
        const int col = gbase + gw * NB + itb;
        if (own) {
          const double *js = Jc + (size_t)col * ld + i0, *zs = Zc + (size_t)col * ld + i0, *ws = Wc + (size_t)col * ld + i0;
          const double2 j01 = *reinterpret_cast<const double2 *>(js), j23 = *reinterpret_cast<const double2 *>(js + 2);
          const double2 z01 = *reinterpret_cast<const double2 *>(zs), z23 = *reinterpret_cast<const double2 *>(zs + 2);
          const double2 w01 = *reinterpret_cast<const double2 *>(ws), w23 = *reinterpret_cast<const double2 *>(ws + 2);
          aa[0] = j01.x; aa[1] = j01.y; aa[2] = j23.x; aa[3] = j23.y;
          bb[0] = z01.x; bb[1] = z01.y; bb[2] = z23.x; bb[3] = z23.y;
          ww[0] = w01.x; ww[1] = w01.y; ww[2] = w23.x; ww[3] = w23.y;
        }
        m = meta[col]; l = logu[col]; g = gsn[col];
      };
      load_col(0, a, b, w, mt, lu, gsnk);
      {   // dx.(invC dx) of the first column; later ones ride along with the previous iteration's reduction
        double part = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) part = fma(a[j] + b[j], w[j], part);
        sd = gsum<32>(part, 0xffffffffu);
      }
      GW_STAMP();   // c0: state loaded
      double *trow_ptr = P.tr.trace + ((size_t)c_local * P.tr.trace_iters + (P.tr.trace_offset + done)) * ld + i0;
      double *lrow_ptr = P.tr.trace_logp + (size_t)c_local * P.tr.trace_iters + (P.tr.trace_offset + done);
      uint32_t *drow_ptr = P.tr.decisions ? P.tr.decisions + (size_t)c_local * P.tr.trace_iters + (P.tr.trace_offset + done) : nullptr;
#pragma unroll 1
      for (int itb = 0; itb < nb; ++itb) {
        // next column (state independent): loads + its dx.(invC dx) partial
        double an[4] = {0, 0, 0, 0}, bn[4] = {0, 0, 0, 0}, wnx[4] = {0, 0, 0, 0};
        uint32_t mtn = 0;
        double lun = 0.0, gn = 0.0, sdp = 0.0;
        if (itb + 1 < nb) {
          load_col(itb + 1, an, bn, wnx, mtn, lun, gn);
#pragma unroll
          for (int j = 0; j < 4; ++j) sdp = fma(an[j] + bn[j], wnx[j], sdp);
        }
        const int run_snooker = (mt >> 8) & 1;
        double prop[4], wn[4], mr, Qn, q_like;
        if (!run_snooker) {
          // prop = q0 + e*gamma*diff + zeta (Dream.py:717); Q(prop) = Q + 2 dx.y + dx.(invC dx)
          double dl[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            prop[j] = x0[j] + a[j] + b[j];
            wn[j] = w[j];
            dl[j] = prop[j] - x0[j];
          }
          double part = fma(dl[1], y0[1], dl[0] * y0[0]) + fma(dl[3], y0[3], dl[2] * y0[2]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            part += __shfl_xor_sync(0xffffffffu, part, o);
            sdp += __shfl_xor_sync(0xffffffffu, sdp, o);
          }
          Qn = (Q0 + (part + part)) + sd;
          q_like = logF - .5 * Qn;
          mr = nan_to_num(q_like) - ntn_last;                                            // Dream.py:334
        } else {
          // snooker_update, Dream.py:827-835 (single-point form); a = z, b = z1 - z2, w = invC z
          const double gamma = gsnk;
          double v[4], t[4];
          double D = 0.0, S = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[j] = (i0 + j < d) ? x0[j] - a[j] : 0.0;
            D = fma(v[j], v[j], D);
            t[j] = b[j] * v[j];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            D += __shfl_xor_sync(0xffffffffu, D, o);
            sdp += __shfl_xor_sync(0xffffffffu, sdp, o);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) S += (D != 0) ? t[j] / D : 0.0;
          const double sc = nan_to_num(gsum<32>(S, 0xffffffffu));
          const double cg = gamma * sc;
          double nn = 0.0, p1 = 0.0, p2 = 0.0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool okd = i0 + j < d;
            const double o = okd ? x0[j] + gamma * (sc * v[j]) : 0.0;
            prop[j] = o;
            const double ww = okd ? o - a[j] : 0.0;
            nn = fma(ww, ww, nn);
            wn[j] = cg * (y0[j] - w[j]);              // invC dx = c (y - invC z)
            const double dl = o - x0[j];
            p1 = fma(dl, y0[j], p1);
            p2 = fma(dl, wn[j], p2);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            nn += __shfl_xor_sync(0xffffffffu, nn, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
            p2 += __shfl_xor_sync(0xffffffffu, p2, o);
          }
          const double norm = sqrt(nn);
          const double snk_logp = (norm != 0 ? log(norm) : 0.0) * (d - 1);
          const double n0 = sqrt(D);
          const double cur = (n0 != 0 ? log(n0) : 0.0) * (d - 1);
          Qn = (Q0 + 2.0 * p1) + p2;
          q_like = logF - .5 * Qn;
          mr = nan_to_num((q_like + snk_logp) - ((1.0 * last_like + last_prior) + cur));   // Dream.py:326-332
        }
        int anydiff = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) anydiff |= (prop[j] != x0[j]);
        const bool accepted = isfinite(mr) && lu < mr;                                   // metrop_select, Dream.py:980-998
        const int changed = accepted && __any_sync(0xffffffffu, anydiff);
        if (changed) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { x0[j] = prop[j]; y0[j] = y0[j] + wn[j]; }
          Q0 = Qn;
          last_prior = 0.0;
          last_like = q_like;
          ntn_last = nan_to_num(q_like);
        }
        if (own) {
          *reinterpret_cast<double2 *>(trow_ptr) = make_double2(x0[0], x0[1]);
          *reinterpret_cast<double2 *>(trow_ptr + 2) = make_double2(x0[2], x0[3]);
          if (P.gw_append && done + itb == P.niter - 1) {   // record_history: the last iteration of the launch
            double *zr = P.st.Z + (size_t)(M + c_global) * ld + i0;
            *reinterpret_cast<double2 *>(zr) = make_double2(x0[0], x0[1]);
            *reinterpret_cast<double2 *>(zr + 2) = make_double2(x0[2], x0[3]);
            for (int pz = 0; pz < P.npeers; ++pz) {   // replicas over NVLink
              double *zp = P.peer_Z[pz] + (size_t)(M + c_global) * ld + i0;
              *reinterpret_cast<double2 *>(zp) = make_double2(x0[0], x0[1]);
              *reinterpret_cast<double2 *>(zp + 2) = make_double2(x0[2], x0[3]);
            }
          }
        }
        if (P.publish_k && P.gw_append && done + itb == P.niter - 1) {
          __threadfence_system();
          __syncwarp();
          if (lane == 0) peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
        }
        if (lane == 0) {
          *lrow_ptr = last_like + last_prior;
          if (drow_ptr)
            *drow_ptr = pack_decision(changed, run_snooker, mt & 15, (mt >> 4) & 15, 1, 0, (mt >> 9) & 1, accepted);
        }
        trow_ptr += ld; lrow_ptr += 1; if (drow_ptr) drow_ptr += 1;
        // rotate in the next column
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = an[j]; b[j] = bn[j]; w[j] = wnx[j]; }
        mt = mtn; lu = lun; gsnk = gn; sd = sdp;
        GW_STAMP();   // c: iteration done
      }
      // park the chain state for the next batch / the epilogue
      if (own) {
        *reinterpret_cast<double2 *>(Xs + cs * ld + i0) = make_double2(x0[0], x0[1]);
        *reinterpret_cast<double2 *>(Xs + cs * ld + i0 + 2) = make_double2(x0[2], x0[3]);
        *reinterpret_cast<double2 *>(Ys + cs * ld + i0) = make_double2(y0[0], y0[1]);
        *reinterpret_cast<double2 *>(Ys + cs * ld + i0 + 2) = make_double2(y0[2], y0[3]);
      }
      if (lane == 0) { cst[cs * 4] = Q0; cst[cs * 4 + 1] = last_prior; cst[cs * 4 + 2] = last_like; }
    }
    GW_STAMP();   // +5: chain of warp 0 advanced
    done += nb;
    group_sync(gid);   // the group's slots are free for the next batch
  }
  GW_STAMP();
  __syncthreads();
  if (warp < nch_cta) {
    const int c_local = cta_chain0 + warp;
    if (own) {
      double *xrow = P.st.X + (size_t)c_local * ld + i0, *yrow = P.st.gauss_Y + (size_t)c_local * ld + i0;
      *reinterpret_cast<double2 *>(xrow) = *reinterpret_cast<const double2 *>(Xs + warp * ld + i0);
      *reinterpret_cast<double2 *>(xrow + 2) = *reinterpret_cast<const double2 *>(Xs + warp * ld + i0 + 2);
      *reinterpret_cast<double2 *>(yrow) = *reinterpret_cast<const double2 *>(Ys + warp * ld + i0);
      *reinterpret_cast<double2 *>(yrow + 2) = *reinterpret_cast<const double2 *>(Ys + warp * ld + i0 + 2);
    }
    if (lane == 0) {
      P.st.gauss_Q[c_local] = cst[warp * 4];
      P.st.last_prior[c_local] = cst[warp * 4 + 1];
      P.st.last_like[c_local] = cst[warp * 4 + 2];
    }
  }
}

// largest batch (iterations) whose tiles fit a group's warps and whose buffers fit shared memory; 0 = kernel not usable
inline int gwin_pick_nb(const dreamzs_config &cfg, int TC, size_t *smem_bytes) {
  if (cfg.ld > 128 || (cfg.ld & 3)) return 0;
  const int gc = gwin_group_chains(TC, 0);
  if (gc * GW_KS > GW_GWARPS) return 0;
  for (int nb = GW_MAXNB; nb >= 1; --nb) {
    if ((gc * nb + GW_GWARPS - 1) / GW_GWARPS > GW_MAXCOLW) continue;
    const GwinLayout L = gwin_layout(cfg.ndim, cfg.ld, TC, nb, cfg.ngamma);
    if (L.bytes <= 227 * 1024) { *smem_bytes = L.bytes; return nb; }
  }
  return 0;
}

template <int TC>
int launch_gwin(StepParams &P, cudaStream_t stream) {
  size_t smem = 0;
  P.gw_nb = gwin_pick_nb(P.cfg, TC, &smem);
  if (P.gw_nb == 0) return DREAMZS_E_UNSUPPORTED;
  P.gw_append = (P.iter_begin + P.niter - 1) % P.cfg.history_thin == 0 ? 1 : 0;
  auto kern = dreamzs_gwin_kernel<TC>;
  static size_t smem_set[64] = {0};
  if (ensure_dynamic_smem(kern, smem, smem_set) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  const int grid = (P.cfg.nchains_local + TC - 1) / TC;
  kern<<<grid, GW_THREADS, smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
