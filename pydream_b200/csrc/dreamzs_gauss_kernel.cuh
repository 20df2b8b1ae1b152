// CTA-synchronous fused step kernel for the dense-Gaussian target (BASELINE config C2), sm_100a.
//
// Why a second kernel: with one warp per chain (dreamzs_step_kernel) every chain streams the whole
// d x d precision matrix out of shared memory once per evaluation, which makes the quadratic form
// shared-memory-bandwidth bound (4 LDS wavefronts per 2 DFMA).  Here a CTA owns TC chains and
// evaluates their quadratic forms TOGETHER: thread tile = 4 dimensions x TC chains, each precision
// entry is loaded once per TC FMAs and x_j is a shared-memory broadcast, so the stage is bound by
// the fp64 pipe instead.
//
//   warp w < TC     owns chain blockIdx.x*TC + w for the element-wise stages (lane g owns the
//                   4-dimension chunk g, exactly as in dreamzs_step_kernel)
//   all 8 warps     share the quadratic forms: thread t -> dimension set iq = t % nq, j-range t / nq
//
// Archive rows are staged by TMA: the row indices of iteration t+2 depend only on the Philox
// counters and on the archive size (constant inside a launch), so one elected lane per chain issues
// cp.async.bulk copies of those rows into a 2-stage shared-memory ring two iterations ahead and the
// gather latency is off the critical path (mbarrier complete_tx).
//
// Arithmetic and RNG consumption are identical to dreamzs_step_kernel / the oracle; only the
// summation order of the quadratic form differs.
#pragma once
#include "dreamzs_step_kernel.cuh"

namespace dreamzs {

constexpr int GK_THREADS = 256;
constexpr int GK_WARPS = GK_THREADS / 32;
constexpr int GK_XS = 10;   // row stride (doubles) of the transposed proposal tile Xs[j][chain]

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct GaussDecisions {
  int run_snooker, cr_idx, delta, lvl_idx;
};

// decisions of one iteration (set_snooker / set_CR / set_DEpair / set_gamma_level, Dream.py:542-599); leaves the
// stream's call counters where astep's generate_proposal_points starts
__device__ __forceinline__ GaussDecisions draw_decisions(const StepParams &P, Stream &s, const double *crp, const double *gp) {
  GaussDecisions dc;
  dc.run_snooker = 0;
  if (P.cfg.snooker != 0) dc.run_snooker = multinomial2(s, P.cfg.snooker) == 0;
  dc.cr_idx = multinomial_index(s, crp, P.cfg.nCR);
  dc.delta = 1;
  if (P.cfg.nDEpairs > 1) {
    const uint4 w = s.block(s.n_randint++, ST_RANDINT, 0);
    dc.delta = 1 + (int)(((uint64_t)w.x * (uint64_t)P.cfg.nDEpairs) >> 32);
  }
  dc.lvl_idx = multinomial_index(s, gp, P.cfg.ngamma);
  return dc;
}

// archive rows of one iteration (sample_from_history): DE -> 2*delta distinct rows, snooker -> z, z1, z2
__device__ __forceinline__ int draw_rows(const Stream &s, const GaussDecisions &dc, int64_t M, int64_t *rows) {
  if (dc.run_snooker) {
    for (int q = 0; q < 3; ++q) {
      const uint4 w = s.block(s.n_sample + q, ST_SAMPLE, 0);
      rows[q] = (int64_t)(((uint64_t)w.x * (uint64_t)M) >> 32);
    }
    return 3;
  }
  const int n = 2 * dc.delta;
  int64_t sorted[2 * DREAMZS_MAX_DEPAIRS];
  uint4 w = make_uint4(0, 0, 0, 0);
  for (int j = 0; j < n; ++j) {
    if ((j & 3) == 0) w = s.block(s.n_sample, ST_SAMPLE, (uint32_t)(j >> 2));
    const uint32_t ww = (j & 3) == 0 ? w.x : (j & 3) == 1 ? w.y : (j & 3) == 2 ? w.z : w.w;
    int64_t rr = (int64_t)(((uint64_t)ww * (uint64_t)(M - j)) >> 32);
    for (int q = 0; q < j; ++q) if (rr >= sorted[q]) rr += 1;
    rows[j] = rr;
    int q = j;
    while (q > 0 && sorted[q - 1] > rr) { sorted[q] = sorted[q - 1]; --q; }
    sorted[q] = rr;
  }
  return n;
}

template <int TC>
__global__ void __launch_bounds__(GK_THREADS, 1) dreamzs_gauss_kernel(const StepParams P) {
  extern __shared__ __align__(16) double smem[];
  const int d = P.cfg.ndim, ld = P.cfg.ld, nq = ld / 4;
  const int NR = max(3, 2 * P.cfg.nDEpairs);   // rows per chain per stage
  // ---- shared memory carve-up
  double *At = smem;                                         // d x ld (precision matrix, transposed)
  double *Xs = At + (size_t)d * ld;                          // ld x GK_XS proposals, transposed [j][chain]
  double *partial = Xs + (size_t)ld * GK_XS;                 // GK_WARPS x 8
  double *rowbuf = partial + GK_WARPS * 8;                   // TC x 2 x NR x ld
  uint64_t *mbar = reinterpret_cast<uint64_t *>(rowbuf + (size_t)TC * 2 * NR * ld);   // TC x 2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double logF = P.st.target_table[0];
  {
    const double *src = P.st.target_table + 2;
    for (int i = tid * 2; i < d * ld; i += GK_THREADS * 2)
      *reinterpret_cast<double2 *>(At + i) = *reinterpret_cast<const double2 *>(src + i);
    for (int i = tid; i < ld * GK_XS; i += GK_THREADS) Xs[i] = 0.0;
    if (tid < TC * 2) mbar_init(mbar + tid, 1);
  }
  __syncthreads();

  const int c_local = blockIdx.x * TC + warp;
  const bool has_chain = warp < TC && c_local < P.cfg.nchains_local;
  const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);
  Ctx<32, 1> c{P, P.st.target_table, nullptr, nullptr, 0xffffffffu, lane, d, ld};
  const int i0 = 4 * lane;
  const bool own = i0 < ld;   // lane owns a chunk

  double x0[1][4] = {{0, 0, 0, 0}};
  double last_prior = 0.0, last_like = 0.0;
  double crp[DREAMZS_MAX_NCR], gp[DREAMZS_MAX_NGAMMA];
  for (int j = 0; j < P.cfg.nCR; ++j) crp[j] = P.st.cr_probs[j];
  for (int j = 0; j < P.cfg.ngamma; ++j) gp[j] = P.st.gamma_probs[j];
  const int64_t M = P.archive_rows;
  const uint32_t row_bytes = (uint32_t)ld * 8u;
  double *myrows = rowbuf + (size_t)warp * 2 * NR * ld;
  uint64_t *mybar = mbar + warp * 2;

  GaussDecisions dcur = {0, 0, 1, 0}, dnext = {0, 0, 1, 0};
  // issue the TMA gathers of iteration `iter` into stage `stage`; returns its decisions
  auto prefetch = [&](int64_t iter, int stage) -> GaussDecisions {
    Stream s; s.init(P.cfg.seed, c_global, (uint32_t)iter);
    GaussDecisions dc = draw_decisions(P, s, crp, gp);
    int64_t rows[2 * DREAMZS_MAX_DEPAIRS];
    const int n = draw_rows(s, dc, M, rows);
    if (lane == 0) {
      fence_proxy_async();   // earlier generic-proxy reads of this stage are ordered before the async writes
      mbar_expect_tx(mybar + stage, row_bytes * (uint32_t)n);
      for (int q = 0; q < n; ++q)
        tma_load_row(myrows + ((size_t)stage * NR + q) * ld, P.st.Z + (size_t)rows[q] * ld, row_bytes, mybar + stage);
    }
    return dc;
  };

  if (has_chain && P.wait_k) {   // sharded archive: the peers' rows of the previous append must have landed
    if (lane == 0) peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error);
    __syncwarp();
  }
  if (has_chain) {
    if (own) {
      const double *xrow = P.st.X + (size_t)c_local * ld + i0;
      const double2 a = *reinterpret_cast<const double2 *>(xrow), b = *reinterpret_cast<const double2 *>(xrow + 2);
      x0[0][0] = a.x; x0[0][1] = a.y; x0[0][2] = b.x; x0[0][3] = b.y;
    }
    last_prior = P.st.last_prior[c_local];
    last_like = P.st.last_like[c_local];
    dcur = prefetch(P.iter_begin, 0);
    if (P.niter > 1) dnext = prefetch(P.iter_begin + 1, 1);
  }

  // quadratic-form tiling
  const int njr = GK_THREADS / nq;
  const int jlen = (d + njr - 1) / njr;
  const int iq = tid % nq, jr = tid / nq;
  const int j0 = min(d, jr * jlen), j1 = (jr < njr) ? min(d, j0 + jlen) : j0;
  const int ia = 2 * iq, ib = 2 * nq + 2 * iq;

#pragma unroll 1
  for (int it = 0; it < P.niter; ++it) {
    const int64_t iter = P.iter_begin + it;
    const int stage = it & 1;
    // ------------------------------------------------------------ stage 1: proposal (warp per chain)
    double prop[1][4] = {{0, 0, 0, 0}};
    double q_prior = 0.0, snk_logp = 0.0, D0 = 0.0;
    bool gamma_one = false;
    Stream s;
    GaussDecisions dc = dcur;
    if (has_chain) {
      s.init(P.cfg.seed, c_global, (uint32_t)iter);
      // the decision draws were made when this iteration was prefetched: advance the call counters past them
      s.n_multinomial = (P.cfg.snooker != 0 ? 1u : 0u) + 2u;
      s.n_randint = P.cfg.nDEpairs > 1 ? 1u : 0u;
      const double CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
      mbar_wait(mybar + stage, (uint32_t)((it >> 1) & 1));
      const double *rb = myrows + (size_t)stage * NR * ld + i0;
      if (!dc.run_snooker) {
        // ---- DE proposal (generate_proposal_points, Dream.py:688-726)
        const int delta = dc.delta;
        double diff[4] = {0, 0, 0, 0};
        if (own) {
          double a[4], b[4];
          for (int j = 0; j < delta; ++j) {
            const double2 a01 = *reinterpret_cast<const double2 *>(rb + (size_t)j * ld), a23 = *reinterpret_cast<const double2 *>(rb + (size_t)j * ld + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(rb + (size_t)(delta + j) * ld), b23 = *reinterpret_cast<const double2 *>(rb + (size_t)(delta + j) * ld + 2);
            if (j == 0) { a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y; b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y; }
            else { a[0] += a01.x; a[1] += a01.y; a[2] += a23.x; a[3] += a23.y; b[0] += b01.x; b[1] += b01.y; b[2] += b23.x; b[3] += b23.y; }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) diff[j] = a[j] - b[j];
        }
        double zeta[4] = {0, 0, 0, 0}, e[4] = {1, 1, 1, 1};
        unsigned reset = 0;
        int dprime = 0;
        if (own && i0 < d) {
          double nz[4];
          normal4(s.block(s.n_normal, ST_NORMAL, (uint32_t)lane), nz);
          const uint4 we = s.block(s.n_uvec, ST_UNIFORM_VEC, (uint32_t)lane);
          const uint4 wu = s.block(s.n_uvec + 1, ST_UNIFORM_VEC, (uint32_t)lane);
          const uint32_t wev[4] = {we.x, we.y, we.z, we.w}, wuv[4] = {wu.x, wu.y, wu.z, wu.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            zeta[j] = 0.0 + P.cfg.zeta * nz[j];
            e[j] = (-P.cfg.lamb + (P.cfg.lamb - (-P.cfg.lamb)) * u32_of(wev[j])) + 1;
            const double U = u32_of(wuv[j]);
            if (i0 + j < d) {
              dprime += (U < CR);
              if (U > CR) reset |= 1u << j;
            }
          }
        }
        s.n_sample += 1; s.n_normal += 1; s.n_uvec += 2;
        dprime = gsum_int<32>(dprime, 0xffffffffu);
        const int unity = multinomial2(s, P.cfg.p_gamma_unity);
        double gamma;
        if (unity == 0) gamma = 1.0;
        else {
          const int di = dprime >= 1 ? dprime - 1 : d - 1;
          gamma = __ldg(P.st.gamma_table + ((size_t)dc.lvl_idx * P.cfg.nDEpairs + (delta - 1)) * d + di);
        }
        gamma_one = gamma == 1.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          double v = x0[0][j] + (e[j] * gamma) * diff[j] + zeta[j];
          if ((reset >> j) & 1u) v = x0[0][j];
          prop[0][j] = (i0 + j < d) ? v : 0.0;
        }
      } else {
        // ---- snooker proposal (snooker_update, Dream.py:798-837, single-point form :827-835)
        (void)multinomial2(s, P.cfg.p_gamma_unity);
        const double gamma = 1.2 + (2.2 - 1.2) * uniform_scalar(s);
        gamma_one = gamma == 1.0;
        s.n_sample += 3;
        double z[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0}, v[4];
        if (own) {
          const double2 z01 = *reinterpret_cast<const double2 *>(rb), z23 = *reinterpret_cast<const double2 *>(rb + 2);
          const double2 a01 = *reinterpret_cast<const double2 *>(rb + ld), a23 = *reinterpret_cast<const double2 *>(rb + ld + 2);
          const double2 b01 = *reinterpret_cast<const double2 *>(rb + 2 * (size_t)ld), b23 = *reinterpret_cast<const double2 *>(rb + 2 * (size_t)ld + 2);
          z[0] = z01.x; z[1] = z01.y; z[2] = z23.x; z[3] = z23.y;
          t[0] = a01.x - b01.x; t[1] = a01.y - b01.y; t[2] = a23.x - b23.x; t[3] = a23.y - b23.y;
        }
        double D = 0.0, S = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = (i0 + j < d) ? x0[0][j] - z[j] : 0.0;
          D = fma(v[j], v[j], D);
          t[j] = t[j] * v[j];
        }
        D = gsum<32>(D, 0xffffffffu);
#pragma unroll
        for (int j = 0; j < 4; ++j) S += (D != 0) ? t[j] / D : 0.0;
        const double sc = nan_to_num(gsum<32>(S, 0xffffffffu));
        double nn = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool okd = i0 + j < d;
          const double o = okd ? x0[0][j] + gamma * (sc * v[j]) : 0.0;
          prop[0][j] = o;
          const double w = okd ? o - z[j] : 0.0;
          nn = fma(w, w, nn);
        }
        const double norm = sqrt(gsum<32>(nn, 0xffffffffu));
        snk_logp = (norm != 0 ? log(norm) : 0.0) * (d - 1);
        D0 = D;
      }
      // the rows of this stage are consumed: refill it with the gathers of iteration it+2
      __syncwarp();
      dcur = dnext;
      if (it + 2 < P.niter) dnext = prefetch(iter + 2, stage);
      if (P.cfg.hardboundaries && !P.all_flat) apply_bounds<32, 1>(c, s, prop);
      // log prior of the proposal (Model.total_logp, model.py:17-28)
      if (!P.all_flat) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = i0 + j;
          if (i < d) {
            const int kind = P.st.prior_kind[i];
            if (kind == DREAMZS_PRIOR_NORMAL) {
              const double b = P.st.prior_b[i], y = (prop[0][j] - P.st.prior_a[i]) / b;
              acc += (-(y * y) / 2.0 - 0.9189385332046727) - log(b);
            } else if (kind == DREAMZS_PRIOR_UNIFORM) {
              const double b = P.st.prior_b[i], y = (prop[0][j] - P.st.prior_a[i]) / b;
              acc += (y >= 0.0 && y <= 1.0) ? 0.0 - log(b) : -INFINITY;
            }
          }
        }
        q_prior = gsum<32>(acc, 0xffffffffu);
      }
      if (own) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Xs[(size_t)(i0 + j) * GK_XS + warp] = prop[0][j];
      }
    }
    __syncthreads();
    // ------------------------------------------------------------ stage 2: TC quadratic forms (whole CTA)
    {
      double y[4][TC];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) y[r][cc] = 0.0;
#pragma unroll 2
      for (int j = j0; j < j1; ++j) {
        const double2 a01 = *reinterpret_cast<const double2 *>(At + (size_t)j * ld + ia);
        const double2 a23 = *reinterpret_cast<const double2 *>(At + (size_t)j * ld + ib);
        double xv[8];
        const double *xr = Xs + (size_t)j * GK_XS;
#pragma unroll
        for (int cp = 0; cp < 4; ++cp) {
          const double2 t2 = *reinterpret_cast<const double2 *>(xr + 2 * cp);
          xv[2 * cp] = t2.x; xv[2 * cp + 1] = t2.y;
        }
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) {
          y[0][cc] = fma(a01.x, xv[cc], y[0][cc]); y[1][cc] = fma(a01.y, xv[cc], y[1][cc]);
          y[2][cc] = fma(a23.x, xv[cc], y[2][cc]); y[3][cc] = fma(a23.y, xv[cc], y[3][cc]);
        }
      }
      double qp[TC];
#pragma unroll
      for (int cc = 0; cc < TC; ++cc) qp[cc] = 0.0;
      if (j1 > j0) {
        const int irow[4] = {ia, ia + 1, ib, ib + 1};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double *xr = Xs + (size_t)irow[r] * GK_XS;
#pragma unroll
          for (int cp = 0; cp < 4; ++cp) {
            const double2 t2 = *reinterpret_cast<const double2 *>(xr + 2 * cp);
            if (2 * cp < TC) qp[2 * cp] = fma(t2.x, y[r][2 * cp], qp[2 * cp]);
            if (2 * cp + 1 < TC) qp[2 * cp + 1] = fma(t2.y, y[r][2 * cp + 1 < TC ? 2 * cp + 1 : 0], qp[2 * cp + 1 < TC ? 2 * cp + 1 : 0]);
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < TC; ++cc) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qp[cc] += __shfl_xor_sync(0xffffffffu, qp[cc], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) partial[warp * 8 + cc] = qp[cc];
      }
    }
    __syncthreads();
    // ------------------------------------------------------------ stage 3: accept, trace, append (warp per chain)
    if (has_chain) {
      double qf = 0.0;
#pragma unroll
      for (int w = 0; w < GK_WARPS; ++w) qf += partial[w * 8 + warp];
      const double q_like = logF - .5 * qf;
      const double last_logp = 1.0 * last_like + last_prior;
      const double q_logp = 1.0 * q_like + q_prior;
      double mr;
      if (dc.run_snooker) {
        const double norm = sqrt(D0);
        const double cur = (norm != 0 ? log(norm) : 0.0) * (d - 1);
        mr = nan_to_num((q_logp + snk_logp) - (last_logp + cur));
      } else mr = nan_to_num(q_logp) - nan_to_num(last_logp);
      bool accepted = false;
      if (isfinite(mr)) accepted = log(uniform_scalar(s)) < mr;
      int changed = 0;
      if (accepted) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { changed |= (prop[0][j] != x0[0][j]); x0[0][j] = prop[0][j]; }
      }
      changed = __any_sync(0xffffffffu, changed);
      if (changed) { last_prior = q_prior; last_like = q_like; }
      const int64_t trow = P.tr.trace_offset + it;
      if (own) {
        double *dst = P.tr.trace + ((size_t)c_local * P.tr.trace_iters + trow) * ld + i0;
        *reinterpret_cast<double2 *>(dst) = make_double2(x0[0][0], x0[0][1]);
        *reinterpret_cast<double2 *>(dst + 2) = make_double2(x0[0][2], x0[0][3]);
        if (iter % P.cfg.history_thin == 0) {   // record_history: only the last iteration of a launch
          double *zr = P.st.Z + (size_t)(M + c_global) * ld + i0;
          *reinterpret_cast<double2 *>(zr) = make_double2(x0[0][0], x0[0][1]);
          *reinterpret_cast<double2 *>(zr + 2) = make_double2(x0[0][2], x0[0][3]);
          for (int pz = 0; pz < P.npeers; ++pz) {   // replicas over NVLink
            double *zp = P.peer_Z[pz] + (size_t)(M + c_global) * ld + i0;
            *reinterpret_cast<double2 *>(zp) = make_double2(x0[0][0], x0[0][1]);
            *reinterpret_cast<double2 *>(zp + 2) = make_double2(x0[0][2], x0[0][3]);
          }
        }
      }
      if (P.publish_k && iter % P.cfg.history_thin == 0) {
        __threadfence_system();
        __syncwarp();
        if (lane == 0) peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
      }
      if (lane == 0) {
        P.tr.trace_logp[(size_t)c_local * P.tr.trace_iters + trow] = last_like + last_prior;
        if (P.tr.decisions)
          P.tr.decisions[(size_t)c_local * P.tr.trace_iters + trow] =
              pack_decision(changed, dc.run_snooker, dc.cr_idx, dc.lvl_idx, dc.delta, 0, gamma_one, accepted);
      }
    }
  }
  if (has_chain) {
    if (own) {
      double *xrow = P.st.X + (size_t)c_local * ld + i0;
      *reinterpret_cast<double2 *>(xrow) = make_double2(x0[0][0], x0[0][1]);
      *reinterpret_cast<double2 *>(xrow + 2) = make_double2(x0[0][2], x0[0][3]);
    }
    if (lane == 0) { P.st.last_prior[c_local] = last_prior; P.st.last_like[c_local] = last_like; }
  }
}

inline size_t gauss_smem_bytes(const dreamzs_config &cfg, int TC) {
  const int NR = 2 * cfg.nDEpairs > 3 ? 2 * cfg.nDEpairs : 3;
  size_t doubles = (size_t)cfg.ndim * cfg.ld + (size_t)cfg.ld * GK_XS + GK_WARPS * 8 + (size_t)TC * 2 * NR * cfg.ld;
  return doubles * sizeof(double) + (size_t)TC * 2 * sizeof(uint64_t);
}

template <int TC>
int launch_gauss(const StepParams &P, cudaStream_t stream) {
  const size_t smem = gauss_smem_bytes(P.cfg, TC);
  auto kern = dreamzs_gauss_kernel<TC>;
  static size_t smem_set[64] = {0};
  if (ensure_dynamic_smem(kern, smem, smem_set) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  const int grid = (P.cfg.nchains_local + TC - 1) / TC;
  kern<<<grid, GK_THREADS, smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
