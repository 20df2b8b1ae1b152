// Fused MT-DREAM(ZS) step kernel for sm_100a.
//
// One lane-group of G lanes owns one chain for all `niter` iterations of the launch; lane g
// owns the 4-dimension chunks g, g+G, ... (32 B of every row -> two 16-B vector accesses per
// lane, consecutive lanes contiguous: archive gathers, trace writes and archive appends are
// fully coalesced).  A Philox block yields exactly the four per-dimension variates of one
// chunk.  Everything of Dream.astep (pydream/Dream.py:193-362) happens in this kernel:
// decisions, archive gather, DE / snooker proposal, crossover, boundary handling, log prior +
// analytic log-likelihood, multi-try selection and reference set, Metropolis accept, trace
// write, archive append.  The archive is read-only during a launch (see dreamzs_step in
// include/dreamzs.h) so no inter-chain synchronisation is needed.
//
// Parity-sensitive element-wise arithmetic follows numpy's evaluation order; the file is
// compiled with -fmad=false and uses explicit fma() only inside reductions (dot products),
// whose summation order differs from numpy's anyway.
#pragma once
#include "dreamzs_common.cuh"
#include "dreamzs_step_params.cuh"

namespace dreamzs {


template <int G, int R>
struct Ctx {
  const StepParams &P;
  const double *table;  // target table (smem or global)
  double *slots;        // this chain's proposal slots (smem), nslots x ld
  double *scal;         // this chain's per-point scalars (smem): prior[k], like[k], snk[k]
  unsigned gmask;
  int g;                // lane within group
  int d, ld;
  __device__ __forceinline__ int dim0(int r) const { return 4 * (g + G * r); }
};

// ---------------------------------------------------------------- log prior + log likelihood
// Model.total_logp (pydream/model.py:17-32).  x: this lane's chunks; xs: the same point in smem.
template <int G, int R>
__device__ __forceinline__ void eval_logp(const Ctx<G, R> &c, const double (&x)[R][4], const double *xs,
                                          double &prior, double &like) {
  const StepParams &P = c.P;
  const int d = c.d;
  prior = 0.0;
  if (!P.all_flat) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = c.dim0(r) + j;
        if (i < d) {
          const int kind = P.st.prior_kind[i];
          if (kind == DREAMZS_PRIOR_NORMAL) {
            const double b = P.st.prior_b[i], y = (x[r][j] - P.st.prior_a[i]) / b;
            acc += (-(y * y) / 2.0 - 0.9189385332046727) - log(b);
          } else if (kind == DREAMZS_PRIOR_UNIFORM) {
            const double b = P.st.prior_b[i], y = (x[r][j] - P.st.prior_a[i]) / b;
            acc += (y >= 0.0 && y <= 1.0) ? 0.0 - log(b) : -INFINITY;
          }
        }
      }
    }
    prior = gsum<G>(acc, c.gmask);
  }
  const double *tb = c.table;
  switch (P.cfg.target_kind) {
    case DREAMZS_TARGET_CONSTANT: like = tb[0]; break;
    case DREAMZS_TARGET_SUMSHIFT: {
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c.dim0(r) + j < d) acc += x[r][j] + tb[0];
      like = gsum<G>(acc, c.gmask);
    } break;
    case DREAMZS_TARGET_GAUSSIAN_DENSE: {
      // y_i = sum_j invC[i][j] x_j.  The device table holds invC TRANSPOSED with row stride ld at
      // offset 2, so the i owned by consecutive lanes are consecutive 16-B reads (conflict-free
      // LDS.128) and x_j is a shared-memory broadcast.  For this product lane g owns the dimension
      // pairs 2g + 2G m (independent of the chunk ownership used by the element-wise stages).
      const double *At = tb + 2;
      double acc = 0.0;
#pragma unroll 1
      for (int i0 = 2 * c.g; i0 < d; i0 += 2 * G) {
        double y0 = 0.0, y1 = 0.0;
        const double *col = At + i0;
#pragma unroll 4
        for (int j = 0; j < d; ++j) {
          const double xj = xs[j];
          const double2 a = *reinterpret_cast<const double2 *>(col + (size_t)j * c.ld);
          y0 = fma(a.x, xj, y0); y1 = fma(a.y, xj, y1);
        }
        const double2 xi = *reinterpret_cast<const double2 *>(xs + i0);   // padded x are 0
        acc = fma(xi.x, y0, acc); acc = fma(xi.y, y1, acc);
      }
      like = tb[0] - .5 * gsum<G>(acc, c.gmask);
    } break;
    case DREAMZS_TARGET_MIXTURE: {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = c.dim0(r) + j;
          if (i < d) {
            const double a = x[r][j] - tb[2 + i], b = x[r][j] - tb[2 + d + i];
            s0 = fma(a, a, s0); s1 = fma(b, b, s1);
          }
        }
      s0 = gsum<G>(s0, c.gmask); s1 = gsum<G>(s1, c.gmask);
      const double l0 = -.5 * s0 + tb[0], l1 = -.5 * s1 + tb[1];
      const double mx = l0 > l1 ? l0 : l1;
      if (fabs(mx) <= DBL_MAX) {       // exp(mx - mx) is exactly 1: one exponential instead of two
        const double e = exp((l0 > l1 ? l1 : l0) - mx);
        like = log(1.0 + e) + mx;
      } else like = log(exp(l0 - mx) + exp(l1 - mx)) + mx;
    } break;
    case DREAMZS_TARGET_BANANA: {
      const double b = tb[0], v1 = tb[1];
      double acc = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = c.dim0(r) + j;
          if (i < d) {
            if (i == 0) acc += (x[r][0] * x[r][0]) / v1;
            else if (i == 1) { const double y2 = x[r][1] + b * (x[r][0] * x[r][0]) - v1 * b; acc += y2 * y2; }
            else acc = fma(x[r][j], x[r][j], acc);
          }
        }
      like = -.5 * gsum<G>(acc, c.gmask);
    } break;
    case DREAMZS_TARGET_EXTERNAL: like = 0.0; break;   // evaluated by the caller between dreamzs_propose and dreamzs_accept
    default: like = nan(""); break;
  }
}

// ---------------------------------------------------------------- boundary handling
// pydream/Dream.py:734-791: reflect once, then redraw uniformly what is still outside (lower
// set first, then upper set; both masks taken before either redraw).
template <int G, int R>
__device__ __forceinline__ void apply_bounds(const Ctx<G, R> &c, Stream &s, double (&p)[R][4]) {
  const StepParams &P = c.P;
  unsigned lo = 0, hi = 0;  // bit (4r+j)
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
  const int any = gsum_int<G>((lo | hi) != 0, c.gmask);
  if (any == 0) return;
  // rare path: ranks of the out-of-bounds dims in dimension order = prefix over chunks (g + G r)
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned m = pass == 0 ? lo : hi;
    const int total = gsum_int<G>(__popc(m), c.gmask);
    if (total == 0) continue;
    const uint32_t call = s.n_rand++;
    int before_round = 0;  // out-of-bounds dims in earlier rounds (all lanes)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int mine = __popc((m >> (4 * r)) & 15u);
      // exclusive prefix over lanes of the group within this round
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
}

// ---------------------------------------------------------------- proposals
struct Decisions { int run_snooker, cr_idx, delta, lvl_idx; double CR; };
struct Bases { uint32_t s, n, u; };  // first call numbers of the current batch (sample / normal / uniform_vec)

__device__ __forceinline__ const double2 *row_ptr(const double *Z, int64_t row, int ld, int i0) {
  return reinterpret_cast<const double2 *>(Z + (size_t)row * ld + i0);
}

// The state-independent part of one DE proposal, generate_proposal_points DE branch (pydream/Dream.py:688-726) with
// sample_from_history (:646-668) and set_gamma (:601-626): the jump J = (e*gamma)*(sum z_r1 - sum z_r2), zeta and the
// crossover mask (bit 4r+j: the dimension keeps the centre value).  p = point index in a batch of n.
template <int G, int R>
__device__ __forceinline__ void de_draw(const Ctx<G, R> &c, Stream &s, const Decisions &dc, const Bases &b, int n,
                                        int p, int64_t M, double (&J)[R][4], double (&zeta)[R][4], unsigned &reset,
                                        bool &gamma_one, const uint4 *w_sample = nullptr, const double *u_unity = nullptr,
                                        float (*nzf)[4] = nullptr) {
  // (w_sample / u_unity: block 0 of the random.sample call and the uniform of the gamma-unity multinomial when the caller
  // has drawn them already -- the two-stage draw kernel makes all scalar draws of an iteration in one lane-parallel pass;
  // nzf: also return the float32 normals zeta was made of, zeta = 0.0 + cfg.zeta * (double)nzf)
  const StepParams &P = c.P;
  const int d = c.d, delta = dc.delta;
  const double *Z = P.st.Z;
  // --- archive rows: random.sample(range(M), 2 delta)
  double diff[R][4];
  if (delta == 1) {
    const uint4 w = w_sample ? *w_sample : s.block(b.s + p, ST_SAMPLE, 0);
    const int64_t r0 = (int64_t)(((uint64_t)w.x * (uint64_t)M) >> 32);
    int64_t r1 = (int64_t)(((uint64_t)w.y * (uint64_t)(M - 1)) >> 32);
    if (r1 >= r0) r1 += 1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = c.dim0(r);
      if (i0 < d) {
        const double2 a01 = __ldg(row_ptr(Z, r0, c.ld, i0)), a23 = __ldg(row_ptr(Z, r0, c.ld, i0) + 1);
        const double2 b01 = __ldg(row_ptr(Z, r1, c.ld, i0)), b23 = __ldg(row_ptr(Z, r1, c.ld, i0) + 1);
        diff[r][0] = a01.x - b01.x; diff[r][1] = a01.y - b01.y; diff[r][2] = a23.x - b23.x; diff[r][3] = a23.y - b23.y;
      } else diff[r][0] = diff[r][1] = diff[r][2] = diff[r][3] = 0.0;
    }
  } else {
    int64_t rows[2 * DREAMZS_MAX_DEPAIRS], sorted[2 * DREAMZS_MAX_DEPAIRS];
    uint4 w = make_uint4(0, 0, 0, 0);
    for (int j = 0; j < 2 * delta; ++j) {
      if ((j & 3) == 0) w = s.block(b.s + p, ST_SAMPLE, (uint32_t)(j >> 2));
      const uint32_t ww = (j & 3) == 0 ? w.x : (j & 3) == 1 ? w.y : (j & 3) == 2 ? w.z : w.w;
      int64_t rr = (int64_t)(((uint64_t)ww * (uint64_t)(M - j)) >> 32);
      for (int q = 0; q < j; ++q) if (rr >= sorted[q]) rr += 1;
      rows[j] = rr;
      int q = j;
      while (q > 0 && sorted[q - 1] > rr) { sorted[q] = sorted[q - 1]; --q; }
      sorted[q] = rr;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = c.dim0(r);
      double a[4] = {0, 0, 0, 0}, bb[4] = {0, 0, 0, 0};
      if (i0 < d) {
        for (int j = 0; j < delta; ++j) {
          const double2 a01 = __ldg(row_ptr(Z, rows[j], c.ld, i0)), a23 = __ldg(row_ptr(Z, rows[j], c.ld, i0) + 1);
          const double2 b01 = __ldg(row_ptr(Z, rows[delta + j], c.ld, i0)), b23 = __ldg(row_ptr(Z, rows[delta + j], c.ld, i0) + 1);
          if (j == 0) { a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y; bb[0] = b01.x; bb[1] = b01.y; bb[2] = b23.x; bb[3] = b23.y; }
          else { a[0] += a01.x; a[1] += a01.y; a[2] += a23.x; a[3] += a23.y; bb[0] += b01.x; bb[1] += b01.y; bb[2] += b23.x; bb[3] += b23.y; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) diff[r][j] = a[j] - bb[j];
    }
  }
  // --- per-dimension variates: zeta (normal), e (uniform), U (uniform); d' = #{U < CR}
  double e[R][4];
  reset = 0;  // bit 4r+j: U > CR  -> dimension keeps the centre value
  int dprime = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint32_t blk = (uint32_t)(c.g + G * r);
    const int i0 = c.dim0(r);
    if (i0 < d) {
      float nf[4];
      normal4f(s.block(b.n + p, ST_NORMAL, blk), nf);
      const double nz[4] = {(double)nf[0], (double)nf[1], (double)nf[2], (double)nf[3]};
      if (nzf) { nzf[r][0] = nf[0]; nzf[r][1] = nf[1]; nzf[r][2] = nf[2]; nzf[r][3] = nf[3]; }
      const uint4 we = s.block(b.u + p, ST_UNIFORM_VEC, blk);
      const uint4 wu = s.block(b.u + n + p, ST_UNIFORM_VEC, blk);
      const uint32_t wev[4] = {we.x, we.y, we.z, we.w}, wuv[4] = {wu.x, wu.y, wu.z, wu.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        zeta[r][j] = 0.0 + P.cfg.zeta * nz[j];
        e[r][j] = (-P.cfg.lamb + (P.cfg.lamb - (-P.cfg.lamb)) * u32_of(wev[j])) + 1;
        const double U = u32_of(wuv[j]);
        if (i0 + j < d) {
          dprime += (U < dc.CR);
          if (U > dc.CR) reset |= 1u << (4 * r + j);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { zeta[r][j] = 0.0; e[r][j] = 1.0; if (nzf) nzf[r][j] = 0.0f; }
    }
  }
  dprime = gsum_int<G>(dprime, c.gmask);
  // --- gamma: the unity draw is always made (Dream.py:615)
  int unity;
  if (u_unity) { unity = (*u_unity < 0.0 + P.cfg.p_gamma_unity) ? 0 : 1; s.n_multinomial++; }
  else unity = multinomial2(s, P.cfg.p_gamma_unity);   // call number advanced by the caller's order
  double gamma;
  if (unity == 0) gamma = 1.0;
  else {
    const int di = dprime >= 1 ? dprime - 1 : d - 1;
    gamma = __ldg(P.st.gamma_table + ((size_t)dc.lvl_idx * P.cfg.nDEpairs + (delta - 1)) * d + di);
  }
  if (gamma == 1.0) gamma_one = true;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) J[r][j] = (e[r][j] * gamma) * diff[r][j];
}

// One DE proposal around `ctr` (Dream.py:717-726)
template <int G, int R>
__device__ __forceinline__ void de_point(const Ctx<G, R> &c, Stream &s, const Decisions &dc, const Bases &b, int n,
                                         int p, int64_t M, const double (&ctr)[R][4], double (&out)[R][4],
                                         bool &gamma_one) {
  double J[R][4], zeta[R][4];
  unsigned reset;
  de_draw<G, R>(c, s, dc, b, n, p, M, J, zeta, reset, gamma_one);
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = c.dim0(r) + j;
      double v = ctr[r][j] + J[r][j] + zeta[r][j];
      if ((reset >> (4 * r + j)) & 1u) v = ctr[r][j];
      out[r][j] = (i < c.d) ? v : 0.0;
    }
}

// The archive rows of one snooker proposal (Dream.py:802-810): z (the projection anchor) and z1 - z2.
template <int G, int R>
__device__ __forceinline__ void snooker_rows(const Ctx<G, R> &c, const Stream &s, const Bases &b, int n, int p, int64_t M,
                                             double (&z)[R][4], double (&t)[R][4], const uint4 *pre = nullptr) {
  const StepParams &P = c.P;
  const int d = c.d;
  const double *Z = P.st.Z;
  // (pre: the three sample blocks when the caller has drawn them already)
  const uint4 wz = pre ? pre[0] : s.block(b.s + p, ST_SAMPLE, 0);
  const uint4 w1 = pre ? pre[1] : s.block(b.s + n + 2 * p, ST_SAMPLE, 0);
  const uint4 w2 = pre ? pre[2] : s.block(b.s + n + 2 * p + 1, ST_SAMPLE, 0);
  const int64_t rz = (int64_t)(((uint64_t)wz.x * (uint64_t)M) >> 32);
  const int64_t r1 = (int64_t)(((uint64_t)w1.x * (uint64_t)M) >> 32);
  const int64_t r2 = (int64_t)(((uint64_t)w2.x * (uint64_t)M) >> 32);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = c.dim0(r);
    if (i0 < d) {
      const double2 z01 = __ldg(row_ptr(Z, rz, c.ld, i0)), z23 = __ldg(row_ptr(Z, rz, c.ld, i0) + 1);
      const double2 a01 = __ldg(row_ptr(Z, r1, c.ld, i0)), a23 = __ldg(row_ptr(Z, r1, c.ld, i0) + 1);
      const double2 b01 = __ldg(row_ptr(Z, r2, c.ld, i0)), b23 = __ldg(row_ptr(Z, r2, c.ld, i0) + 1);
      z[r][0] = z01.x; z[r][1] = z01.y; z[r][2] = z23.x; z[r][3] = z23.y;
      t[r][0] = a01.x - b01.x; t[r][1] = a01.y - b01.y; t[r][2] = a23.x - b23.x; t[r][3] = a23.y - b23.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { z[r][j] = 0.0; t[r][j] = 0.0; }
    }
  }
}

// snooker_update (pydream/Dream.py:811-837) given the rows: the proposal, snooker_logp of the point and D = |ctr - z|^2
// (t = z1 - z2 is overwritten).
template <int G, int R>
__device__ __forceinline__ void snooker_compute(const Ctx<G, R> &c, int n, double gamma, const double (&ctr)[R][4],
                                                const double (&z)[R][4], double (&t)[R][4], double (&out)[R][4],
                                                double &snk_logp, double &Dout) {
  const int d = c.d;
  double v[R][4];
  double D = 0.0, S = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = c.dim0(r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = i0 + j < d;
      v[r][j] = ok ? ctr[r][j] - z[r][j] : 0.0;
      D = fma(v[r][j], v[r][j], D);
      t[r][j] = t[r][j] * v[r][j];   // (z1 - z2) * (q0 - z)
    }
  }
  D = gsum<G>(D, c.gmask);
  double sc;
  if (n > 1) {           // Dream.py:816-822: sum(...)/D, nan_to_num on the product with v
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) S += t[r][j];
    sc = gsum<G>(S, c.gmask) / D;
  } else {               // Dream.py:827-833: element-wise divide (0 where D == 0), nan_to_num on the sum
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) S += (D != 0) ? t[r][j] / D : 0.0;
    sc = nan_to_num(gsum<G>(S, c.gmask));
  }
  double nn = 0.0;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool ok = c.dim0(r) + j < d;
      const double zp = (n > 1) ? nan_to_num(sc * v[r][j]) : sc * v[r][j];
      const double o = ok ? ctr[r][j] + gamma * zp : 0.0;
      out[r][j] = o;
      const double w = ok ? o - z[r][j] : 0.0;
      nn = fma(w, w, nn);
    }
  const double norm = sqrt(gsum<G>(nn, c.gmask));
  snk_logp = (norm != 0 ? log(norm) : 0.0) * (d - 1);   // log(where=False) pinned to 0
  Dout = D;
}

// One snooker proposal, snooker_update (pydream/Dream.py:798-837); gamma was drawn once per batch.
// Returns snooker_logp of the point and D = |ctr - z|^2.
template <int G, int R>
__device__ __forceinline__ void snooker_point(const Ctx<G, R> &c, Stream &s, const Bases &b, int n, int p, int64_t M,
                                              double gamma, const double (&ctr)[R][4], double (&out)[R][4],
                                              double &snk_logp, double &Dout) {
  double z[R][4], t[R][4];
  snooker_rows<G, R>(c, s, b, n, p, M, z, t);
  snooker_compute<G, R>(c, n, gamma, ctr, z, t, out, snk_logp, Dout);
}

// The call numbers a batch of n points consumes (the bookkeeping at the end of gen_eval_batch), without the rand()
// redraws of the boundary handling, which are data dependent and carried separately.
__device__ __forceinline__ void advance_batch_counters(Stream &s, const Decisions &dc, int n) {
  if (dc.run_snooker) { s.n_multinomial += 1; s.n_uscal += 1; s.n_sample += 3 * n; }
  else { s.n_sample += n; s.n_normal += n; s.n_uvec += 2 * n; s.n_multinomial += n; }
}

template <int G, int R>
__device__ __forceinline__ void store_slot(const Ctx<G, R> &c, double *slot, const double (&x)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = c.dim0(r);
    if (i0 < c.ld) {
      *reinterpret_cast<double2 *>(slot + i0) = make_double2(x[r][0], x[r][1]);
      *reinterpret_cast<double2 *>(slot + i0 + 2) = make_double2(x[r][2], x[r][3]);
    }
  }
}
template <int G, int R>
__device__ __forceinline__ void load_slot(const Ctx<G, R> &c, const double *slot, double (&x)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = c.dim0(r);
    if (i0 < c.ld) {
      const double2 a = *reinterpret_cast<const double2 *>(slot + i0), b = *reinterpret_cast<const double2 *>(slot + i0 + 2);
      x[r][0] = a.x; x[r][1] = a.y; x[r][2] = b.x; x[r][3] = b.y;
    } else x[r][0] = x[r][1] = x[r][2] = x[r][3] = 0.0;
  }
}
template <int G, int R>
__device__ __forceinline__ void store_row(const Ctx<G, R> &c, double *row, const double (&x)[R][4]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i0 = c.dim0(r);
    if (i0 < c.ld) {
      *reinterpret_cast<double2 *>(row + i0) = make_double2(x[r][0], x[r][1]);
      *reinterpret_cast<double2 *>(row + i0 + 2) = make_double2(x[r][2], x[r][3]);
    }
  }
}

// A batch of n points around `ctr` (generate_proposal_points): point p lands in slot `slot0+p`
// (or always slot0 when eval_now), its scalars in pri/lik/snk arrays at index p.
// Returns gamma_one for the batch (np.any(gamma == 1.0)).
template <int G, int R>
__device__ __forceinline__ bool gen_eval_batch(const Ctx<G, R> &c, Stream &s, const Decisions &dc, int n, int64_t M,
                                               const double (&ctr)[R][4], int slot0, bool one_slot, double *pri,
                                               double *lik, double *snk, double &D0) {
  const StepParams &P = c.P;
  Bases b = {s.n_sample, s.n_normal, s.n_uvec};
  bool gamma_one = false;
  double gamma = 0.0;
  uint32_t m_base = s.n_multinomial;
  if (dc.run_snooker) {
    (void)multinomial2(s, P.cfg.p_gamma_unity);              // drawn and discarded (Dream.py:615-618)
    gamma = 1.2 + (2.2 - 1.2) * uniform_scalar(s);
    if (gamma == 1.0) gamma_one = true;
  }
  for (int p = 0; p < n; ++p) {
    double pt[R][4];
    double sl = 0.0;
    if (dc.run_snooker) {
      double D;
      snooker_point<G, R>(c, s, b, n, p, M, gamma, ctr, pt, sl, D);
      if (p == 0) D0 = D;
    } else {
      s.n_multinomial = m_base + p;                          // gamma-unity draw of point p
      de_point<G, R>(c, s, dc, b, n, p, M, ctr, pt, gamma_one);
    }
    if (P.cfg.hardboundaries && !P.all_flat) apply_bounds<G, R>(c, s, pt);
    double *slot = c.slots + (size_t)(one_slot ? slot0 : slot0 + p) * c.ld;
    __syncwarp(c.gmask);
    store_slot<G, R>(c, slot, pt);
    __syncwarp(c.gmask);
    double pr, lk;
    eval_logp<G, R>(c, pt, slot, pr, lk);
    if (c.g == 0) { pri[p] = pr; lik[p] = lk; snk[p] = sl; }
  }
  // call numbers consumed by the batch
  if (dc.run_snooker) s.n_sample = b.s + 3 * n;
  else { s.n_sample = b.s + n; s.n_normal = b.n + n; s.n_uvec = b.u + 2 * n; s.n_multinomial = m_base + n; }
  __syncwarp(c.gmask);
  return gamma_one;
}

// The multi-try path calls the batch generator twice (proposals, reference set): ONE out-of-line copy keeps the
// kernel's instruction footprint down (the fused multi-try kernel is instruction-fetch bound: ncu no_instruction
// stalls 3.5 warps per issue at d=10); the single-try path keeps the inlined copy, whose state stays in registers.
template <int G, int R>
__device__ __noinline__ bool gen_eval_batch_mt(const Ctx<G, R> &c, Stream &s, const Decisions &dc, int n, int64_t M,
                                               const double (&ctr)[R][4], int slot0, bool one_slot, double *pri,
                                               double *lik, double *snk, double &D0) {
  return gen_eval_batch<G, R>(c, s, dc, n, M, ctr, slot0, one_slot, pri, lik, snk, D0);
}

// MT = the multi-try code is compiled in (a separate instantiation: its out-of-line batch generator takes the chain
// state by reference, which would push the single-try kernel's registers into local memory as well)
template <int G, int R, bool MT>
__global__ void __launch_bounds__(128) dreamzs_step_kernel(const StepParams P) {
  extern __shared__ __align__(16) double smem[];
  const int d = P.cfg.ndim, ld = P.cfg.ld, k = P.cfg.multitry;
  // ---- stage the target table (dense precision matrix) in shared memory
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
  const int per_chain = P.nslots * ld + 3 * DREAMZS_MAX_MULTITRY;
  Ctx<G, R> c{P, table, sm_chain + (size_t)chain_in_cta * per_chain, nullptr,
              G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1))), lane & (G - 1), d, ld};
  c.scal = c.slots + (size_t)P.nslots * ld;
  double *pri = c.scal, *lik = c.scal + DREAMZS_MAX_MULTITRY, *snk = c.scal + 2 * DREAMZS_MAX_MULTITRY;
  const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);

  double x0[R][4];
  {
    const double *xrow = P.st.X + (size_t)c_local * ld;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i0 = c.dim0(r);
      if (i0 < ld) {
        const double2 a = *reinterpret_cast<const double2 *>(xrow + i0), b = *reinterpret_cast<const double2 *>(xrow + i0 + 2);
        x0[r][0] = a.x; x0[r][1] = a.y; x0[r][2] = b.x; x0[r][3] = b.y;
      } else x0[r][0] = x0[r][1] = x0[r][2] = x0[r][3] = 0.0;
    }
  }
  if (P.init_only) {   // first-call branch of astep, Dream.py:266-268
    double pr, lk;
    store_slot<G, R>(c, c.slots, x0);
    __syncwarp(c.gmask);
    eval_logp<G, R>(c, x0, c.slots, pr, lk);
    if (c.g == 0) { P.st.last_prior[c_local] = pr; P.st.last_like[c_local] = lk; }
    return;
  }
  double last_prior = P.st.last_prior[c_local], last_like = P.st.last_like[c_local];
  // temperature of astep(q0, T, ...): log_ps = T * log_like + log_prior everywhere (Dream.py:243, 268, 274, 279, 303, 899)
  const double Tc = P.temperature ? P.temperature[c_local] : 1.0;

  double crp[DREAMZS_MAX_NCR], gp[DREAMZS_MAX_NGAMMA];
  for (int j = 0; j < P.cfg.nCR; ++j) crp[j] = P.st.cr_probs[j];
  for (int j = 0; j < P.cfg.ngamma; ++j) gp[j] = P.st.gamma_probs[j];

  const int64_t M = P.archive_rows;
  if (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) return;   // an earlier wait timed out: the host raises
  if (P.wait_k) {   // sharded archive: the peers' rows of the previous append must have landed in this replica
    int ok = 1;
    if (c.g == 0) ok = peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error) ? 1 : 0;
    ok = __shfl_sync(c.gmask, ok, lane & ~(G - 1));
    if (!ok) return;   // timed out: nothing is sampled from a replica that may miss rows; states stay as they are
  }
#pragma unroll 1
  for (int it = 0; it < P.niter; ++it) {
    const int64_t iter = P.iter_begin + it;
    Stream s; s.init(P.cfg.seed, c_global, (uint32_t)iter);
    Decisions dc; dc.run_snooker = 0;
    if (P.cfg.snooker != 0) dc.run_snooker = multinomial2(s, P.cfg.snooker) == 0;      // set_snooker, Dream.py:542-554
    dc.cr_idx = multinomial_index(s, crp, P.cfg.nCR);                                  // set_CR, :556-569
    dc.CR = (double)(dc.cr_idx + 1) / (double)P.cfg.nCR;
    dc.delta = 1;
    if (P.cfg.nDEpairs > 1) {                                                          // set_DEpair, :571-583
      const uint4 w = s.block(s.n_randint++, ST_RANDINT, 0);
      dc.delta = 1 + (int)(((uint64_t)w.x * (uint64_t)P.cfg.nDEpairs) >> 32);
    }
    dc.lvl_idx = multinomial_index(s, gp, P.cfg.ngamma);                               // set_gamma_level, :585-599

    const double last_logp = Tc * last_like + last_prior;
    double D0 = 0.0;
    bool accepted = false, gamma_one;
    int sel = 0;
    double new_prior, new_like;
    double q[R][4];
    if (!MT || k == 1) {
      double q_prior, q_like, snk0;
      if (P.ext_phase == 2) {
        // split step, second half: the proposal, its log prior and snooker terms come from dreamzs_propose, its
        // log-likelihood from the caller; the random stream resumes where generate_proposal_points left it
        const double *ax = P.ext_aux + (size_t)c_local * 4;
        q_prior = ax[0]; snk0 = ax[1]; D0 = ax[2]; gamma_one = ax[3] != 0.0;
        q_like = P.ext_like[c_local];
        const double *prow = P.ext_prop + (size_t)c_local * ld;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i0 = c.dim0(r);
          if (i0 < ld) {
            const double2 a = *reinterpret_cast<const double2 *>(prow + i0), b = *reinterpret_cast<const double2 *>(prow + i0 + 2);
            q[r][0] = a.x; q[r][1] = a.y; q[r][2] = b.x; q[r][3] = b.y;
          } else q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0.0;
        }
        s.n_uscal = dc.run_snooker ? 1u : 0u;
      } else {
        gamma_one = gen_eval_batch<G, R>(c, s, dc, 1, M, x0, 0, true, pri, lik, snk, D0);
        q_prior = pri[0]; q_like = lik[0]; snk0 = snk[0];
        load_slot<G, R>(c, c.slots, q);
        if (P.ext_phase == 1) {   // split step, first half: hand the proposal to the caller; no state changes
          store_row<G, R>(c, P.ext_prop + (size_t)c_local * ld, q);
          if (c.g == 0) {
            double *ax = P.ext_aux + (size_t)c_local * 4;
            ax[0] = q_prior; ax[1] = snk0; ax[2] = D0; ax[3] = gamma_one ? 1.0 : 0.0;
          }
          return;
        }
      }
      const double q_logp = Tc * q_like + q_prior;
      double mr;
      if (dc.run_snooker) {                                                            // Dream.py:326-332
        const double norm = sqrt(D0);
        const double cur = (norm != 0 ? log(norm) : 0.0) * (d - 1);
        mr = nan_to_num((q_logp + snk0) - (last_logp + cur));
      } else mr = nan_to_num(q_logp) - nan_to_num(last_logp);                          // Dream.py:334
      if (isfinite(mr)) accepted = log(uniform_scalar(s)) < mr;                        // metrop_select, :980-998
      new_prior = q_prior; new_like = q_like;
    } else if constexpr (MT) {
      double *rpri = pri + k, *rlik = lik + k, *rsnk = snk + k;   // MAX_MULTITRY >= 2k is not required: see host check
      // Split step with a caller-evaluated likelihood (ext_phase 1 / 2 / 3 = propose / select / accept): the
      // 2k-1 points live in ext_prop [chain][2k-1][ld] (proposals first, then the reference set), their
      // log-likelihoods in ext_like [chain][2k-1], and the scalars the next phase needs in ext_aux [chain][4k+4]:
      //   [0,k) log prior, [k,2k) snooker logp of the proposals, 2k: rand() calls made so far, 2k+1: selected index,
      //   [2k+2,3k+1) / [3k+1,4k) the same of the reference points, 4k: rand() calls, 4k+1: gamma == 1 flag,
      //   4k+2: proposal batches regenerated so far (phase 4 = dreamzs_repropose, the loop of Dream.py:278-289)
      const int npts = 2 * k - 1;
      double *ax = P.ext_phase ? P.ext_aux + (size_t)c_local * (4 * k + 4) : nullptr;
      double *xprop = P.ext_phase ? P.ext_prop + (size_t)c_local * npts * ld : nullptr;
      const double *xlike = P.ext_phase >= 2 ? P.ext_like + (size_t)c_local * npts : nullptr;
      if (P.ext_phase <= 1) {
        for (int guard = 0;; ++guard) {                                                // Dream.py:278-289
          gamma_one = gen_eval_batch_mt<G, R>(c, s, dc, k, M, x0, 0, false, pri, lik, snk, D0);
          if (P.ext_phase == 1) break;
          bool anyfinite = false;
          for (int p = 0; p < k; ++p) anyfinite |= isfinite(Tc * lik[p] + pri[p]);
          if (anyfinite || guard >= 1000) break;
        }
        if (P.ext_phase == 1) {   // hand the k proposals to the caller
          for (int p = 0; p < k; ++p) {
            load_slot<G, R>(c, c.slots + (size_t)p * ld, q);
            store_row<G, R>(c, xprop + (size_t)p * ld, q);
          }
          if (c.g == 0) {
            for (int p = 0; p < k; ++p) { ax[p] = pri[p]; ax[k + p] = snk[p]; }
            ax[2 * k] = (double)s.n_rand;
            ax[4 * k + 2] = 0.0;
          }
          return;
        }
      } else {                    // the proposals' scalars come back; the stream resumes after their draws
        if (c.g == 0)
          for (int p = 0; p < k; ++p) { pri[p] = ax[p]; snk[p] = ax[k + p]; lik[p] = xlike[p]; }
        __syncwarp(c.gmask);
        const int rounds = (int)ax[4 * k + 2];          // batches regenerated so far: each consumed a batch's draws
        for (int r = 0; r <= rounds; ++r) advance_batch_counters(s, dc, k);
        s.n_rand = (uint32_t)ax[2 * k];
        bool anyfinite = false;
        for (int p = 0; p < k; ++p) anyfinite |= isfinite(Tc * lik[p] + pri[p]);
        if (P.ext_phase == 4) {   // dreamzs_repropose: a chain whose k proposals are all non-finite draws the next batch
          if (anyfinite) return;
          gamma_one = gen_eval_batch_mt<G, R>(c, s, dc, k, M, x0, 0, false, pri, lik, snk, D0);
          for (int p = 0; p < k; ++p) {
            load_slot<G, R>(c, c.slots + (size_t)p * ld, q);
            store_row<G, R>(c, xprop + (size_t)p * ld, q);
          }
          if (c.g == 0) {
            for (int p = 0; p < k; ++p) { ax[p] = pri[p]; ax[k + p] = snk[p]; }
            ax[2 * k] = (double)s.n_rand;
            ax[4 * k + 2] = (double)(rounds + 1);
            atomicAdd(P.ext_error, 1);              // (here: the number of chains that drew a new batch)
          }
          return;
        }
        if (P.ext_phase == 2 && !anyfinite && c.g == 0) atomicExch(P.ext_error, 1);   // the caller gave up regenerating
      }
      if (P.ext_phase <= 2) {
        // mt_choose_proposal_pt, Dream.py:883-917
        double mx = Tc * lik[0] + pri[0];
        for (int p = 1; p < k; ++p) { const double v = Tc * lik[p] + pri[p]; if (v > mx) mx = v; }
        double prob[DREAMZS_MAX_MULTITRY], sum = 0.0;
        for (int p = 0; p < k; ++p) { prob[p] = exp((Tc * lik[p] + pri[p]) - mx); sum = (p == 0) ? prob[0] : sum + prob[p]; }
        for (int p = 0; p < k; ++p) prob[p] = prob[p] / sum;
        sel = multinomial_index(s, prob, k);
      } else {
        sel = (int)ax[2 * k + 1];
        s.n_multinomial += 1;
      }
      if (P.ext_phase == 0) load_slot<G, R>(c, c.slots + (size_t)sel * ld, q);
      else {
        const double *prow = xprop + (size_t)sel * ld;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i0 = c.dim0(r);
          if (i0 < ld) {
            const double2 a = *reinterpret_cast<const double2 *>(prow + i0), b = *reinterpret_cast<const double2 *>(prow + i0 + 2);
            q[r][0] = a.x; q[r][1] = a.y; q[r][2] = b.x; q[r][3] = b.y;
          } else q[r][0] = q[r][1] = q[r][2] = q[r][3] = 0.0;
        }
      }
      if (P.ext_phase <= 2) {
        // reference set around the selected proposal, Dream.py:295-303
        gamma_one = gen_eval_batch_mt<G, R>(c, s, dc, k - 1, M, q, k, P.ext_phase == 0, rpri, rlik, rsnk, D0);
        if (P.ext_phase == 2) {   // hand the k-1 reference points to the caller
          double rq[R][4];
          for (int p = 0; p < k - 1; ++p) {
            load_slot<G, R>(c, c.slots + (size_t)(k + p) * ld, rq);
            store_row<G, R>(c, xprop + (size_t)(k + p) * ld, rq);
          }
          if (c.g == 0) {
            ax[2 * k + 1] = (double)sel;
            for (int p = 0; p < k - 1; ++p) { ax[2 * k + 2 + p] = rpri[p]; ax[3 * k + 1 + p] = rsnk[p]; }
            ax[4 * k] = (double)s.n_rand;
            ax[4 * k + 1] = gamma_one ? 1.0 : 0.0;
          }
          return;
        }
      } else {
        if (c.g == 0)
          for (int p = 0; p < k - 1; ++p) { rpri[p] = ax[2 * k + 2 + p]; rsnk[p] = ax[3 * k + 1 + p]; rlik[p] = xlike[k + p]; }
        __syncwarp(c.gmask);
        advance_batch_counters(s, dc, k - 1);
        s.n_rand = (uint32_t)ax[4 * k];
        gamma_one = ax[4 * k + 1] != 0.0;
      }
      double tp[DREAMZS_MAX_MULTITRY], trf[DREAMZS_MAX_MULTITRY];
      double m2 = -INFINITY;
      for (int p = 0; p < k; ++p) {
        const double lps = Tc * lik[p] + pri[p];
        const double rl = (p == k - 1) ? last_like : rlik[p], rp = (p == k - 1) ? last_prior : rpri[p];
        const double rlps = Tc * rl + rp;
        if (dc.run_snooker) {                                                          // Dream.py:306-313
          const double rs = (p == k - 1) ? 0.0 : rsnk[p];
          tp[p] = lps + snk[p]; trf[p] = rlps + rs + snk[p];
        } else { tp[p] = lps; trf[p] = rlps; }
        if (p == 0) m2 = tp[0];
        if (tp[p] > m2) m2 = tp[p];
        if (trf[p] > m2) m2 = trf[p];
      }
      double swp = 0.0, swr = 0.0;
      for (int p = 0; p < k; ++p) {
        const double a = exp(tp[p] - m2), b = exp(trf[p] - m2);
        swp = p == 0 ? a : swp + a; swr = p == 0 ? b : swr + b;
      }
      const double mr = nan_to_num(log(swp / swr));                                    // Dream.py:320-323
      if (isfinite(mr)) accepted = log(uniform_scalar(s)) < mr;
      new_prior = pri[sel]; new_like = lik[sel];
    }
    // state update (Dream.py:336-347: "accepted" is inferred from the state having changed)
    int changed = 0;
    if (accepted) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) { changed |= (q[r][j] != x0[r][j]); x0[r][j] = q[r][j]; }
    }
    changed = gsum_int<G>(changed, c.gmask) != 0;
    if (changed) { last_prior = new_prior; last_like = new_like; }
    // trace (core.py:114-115)
    const int64_t trow = P.tr.trace_offset + it;
    store_row<G, R>(c, P.tr.trace + ((size_t)c_local * P.tr.trace_iters + trow) * ld, x0);
    if (c.g == 0) {
      P.tr.trace_logp[(size_t)c_local * P.tr.trace_iters + trow] = Tc * last_like + last_prior;   // core.py:115 (T = 1), :176
      if (P.tr.decisions)
        P.tr.decisions[(size_t)c_local * P.tr.trace_iters + trow] =
            pack_decision(changed, dc.run_snooker, dc.cr_idx, dc.lvl_idx, dc.delta, sel, gamma_one, accepted);
    }
    // record_history (Dream.py:360-362, 919-938): only the last iteration of a launch may append
    if (iter % P.cfg.history_thin == 0) {
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


// host-side launcher of one <G, R> instantiation (defined in dreamzs_step_inst.cu, one object per variant)
template <int G, int R>
int launch_step(const StepParams &P, int threads, size_t smem, cudaStream_t stream) {
  const int chains_per_cta = (threads / 32) * (32 / G);
  const int grid = (P.cfg.nchains_local + chains_per_cta - 1) / chains_per_cta;
  auto kern = P.cfg.multitry > 1 ? dreamzs_step_kernel<G, R, true> : dreamzs_step_kernel<G, R, false>;
  if (smem > 48 * 1024) {
    static size_t smem_set[2][64] = {{0}};     // one cache per instantiation (single-try, multi-try)
    if (ensure_dynamic_smem(kern, smem, smem_set[P.cfg.multitry > 1 ? 1 : 0]) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  }
  kern<<<grid, threads, smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
