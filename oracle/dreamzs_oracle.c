/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the MT-DREAM(ZS) step path of
 * LoLab-MSM/PyDREAM in plain C.  Only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline / --impl reference legs may link or call it.  The product (pydream_b200,
 * libdreamzs.so) never does.
 *
 * Parity status: PINNED.  This file is checked (tests/test_oracle_golden.py) against
 * outputs of the UNMODIFIED reference executed in lock-step with injected counter-based
 * RNG (oracle/ref_harness.py; vectors in tests/golden/, generator tests/golden/make_golden.py)
 * and against the known-answer values the reference's own tests hold
 * (gamma table, pydream/tests/test_dream.py:68-76).
 *
 * Each function cites the reference code it restates (paths relative to the reference root).
 * Scheduling semantics are the synchronous ones documented in DESIGN.md: all chains of
 * iteration t read the archive as it stood after iteration t-1; record_history and the
 * adaptation updates are applied after the sweep in chain order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include "../include/dreamzs.h"

/* ------------------------------------------------------------------ RNG contract */
enum { ST_MULTINOMIAL = 0, ST_SAMPLE, ST_NORMAL, ST_UNIFORM_VEC, ST_UNIFORM_SCAL, ST_RAND, ST_RANDINT };

typedef struct { uint32_t k0, k1, chain, iter; uint32_t calls[8]; } stream_t;

/* Philox4x32-10, Salmon et al. SC'11 (Random123 constants). */
static void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void stream_init(stream_t *s, uint64_t seed, uint32_t chain, uint32_t iter) {
  s->k0 = (uint32_t)seed; s->k1 = (uint32_t)(seed >> 32); s->chain = chain; s->iter = iter;
  memset(s->calls, 0, sizeof(s->calls));
}
static uint32_t stream_next(stream_t *s, int st) { uint32_t n = s->calls[st]++; return (n << 3) | (uint32_t)st; }
static void stream_block(const stream_t *s, uint32_t c1, uint32_t block, uint32_t w[4]) {
  philox4x32(block, c1, s->iter, s->chain, s->k0, s->k1, w);
}
static double u53_of(uint32_t w0, uint32_t w1) {
  return (double)(((uint64_t)(w0 >> 5) << 26) + (uint64_t)(w1 >> 6)) * (1.0 / 9007199254740992.0);
}
static double stream_uniform53(stream_t *s, int st) {
  uint32_t w[4]; stream_block(s, stream_next(s, st), 0, w); return u53_of(w[0], w[1]);
}
/* n 32-bit uniforms w*2^-32 */
static void stream_uniform32_vec(stream_t *s, int st, int n, double *out) {
  uint32_t c1 = stream_next(s, st), w[4];
  for (int b = 0; 4 * b < n; ++b) {
    stream_block(s, c1, (uint32_t)b, w);
    for (int j = 0; j < 4 && 4 * b + j < n; ++j) out[4 * b + j] = (double)w[j] * (1.0 / 4294967296.0);
  }
}
/* Normal variates of the contract (stream 2): Box-Muller in float32 on 24-bit uniforms, written with operations
 * that round identically everywhere (float add / multiply / fmaf / sqrtf on fixed Cephes coefficients); the recipe
 * is documented in oracle/philox.py (normal_pairs32).  Compiled with -ffp-contract=off: only the explicit fmaf()
 * calls fuse. */
static void normal_pair32(uint32_t w0, uint32_t w1, float *n0, float *n1) {
  static const float LOG_P[9] = {7.0376836292E-2f, -1.1514610310E-1f, 1.1676998740E-1f, -1.2420140846E-1f, 1.4249322787E-1f,
                                 -1.6668057665E-1f, 2.0000714765E-1f, -2.4999993993E-1f, 3.3333331174E-1f};
  const float a = (float)((w0 >> 8) + 1u);
  uint32_t bits; memcpy(&bits, &a, 4);
  int E = (int)(bits >> 23) - 127;
  uint32_t mb = (bits & 0x007FFFFFu) | 0x3F800000u;
  float m; memcpy(&m, &mb, 4);
  if (m > 1.41421356f) { m = m * 0.5f; E += 1; }
  const float t = m - 1.0f, z = t * t;
  float P = LOG_P[0];
  for (int i = 1; i < 9; ++i) P = fmaf(P, t, LOG_P[i]);
  float y = t * (z * P);
  y = fmaf(-0.5f, z, y);
  const float logm = t + y;
  const float L = fmaf((float)(E - 24), 0.6931471805599453f, logm);
  const float val = -2.0f * L;
  const float r = sqrtf(val > 0.0f ? val : 0.0f);
  const int32_t w24 = (int32_t)(w1 >> 8);
  const int32_t k = (w24 + (1 << 21)) >> 22;
  const float g = (float)(w24 - (k << 22)) * 5.9604644775390625e-08f;   /* 2^-24, exact */
  const float phi = g * 6.283185307179586f;
  const float zz = phi * phi;
  const float sp = fmaf(fmaf(-1.9515295891E-4f, zz, 8.3321608736E-3f), zz, -1.6666654611E-1f);
  const float cp = fmaf(fmaf(2.443315711809948E-5f, zz, -1.388731625493765E-3f), zz, 4.166664568298827E-2f);
  const float s = fmaf(phi * zz, sp, phi);
  const float c = fmaf(zz * zz, cp, fmaf(-0.5f, zz, 1.0f));
  const int q = k & 3;
  const float cs = q == 0 ? c : q == 1 ? -s : q == 2 ? -c : s;
  const float sn = q == 0 ? s : q == 1 ? c : q == 2 ? -s : -c;
  *n0 = r * cs; *n1 = r * sn;
}
static void stream_normal_vec(stream_t *s, int n, double *out) {
  uint32_t c1 = stream_next(s, ST_NORMAL), w[4];
  for (int b = 0; 4 * b < n; ++b) {
    stream_block(s, c1, (uint32_t)b, w);
    for (int h = 0; h < 2; ++h) {
      float n0, n1;
      normal_pair32(w[2 * h], w[2 * h + 1], &n0, &n1);
      int i = 4 * b + 2 * h;
      if (i < n) out[i] = (double)n0;
      if (i + 1 < n) out[i + 1] = (double)n1;
    }
  }
}
/* test hook: the float32 normals of word pairs (tests/test_philox.py pins the three implementations to each other) */
void dreamzs_oracle_normal_pairs32(const uint32_t *w0, const uint32_t *w1, int64_t n, float *n0, float *n1) {
  for (int64_t i = 0; i < n; ++i) normal_pair32(w0[i], w1[i], n0 + i, n1 + i);
}
/* test hook: fmaf, to check the numpy emulation in oracle/philox.py */
void dreamzs_oracle_fmaf(const float *a, const float *b, const float *c, int64_t n, float *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = fmaf(a[i], b[i], c[i]);
}
/* random.sample(range(M), n) restated on the counter stream (pydream/Dream.py:662-664) */
static void stream_sample(stream_t *s, int64_t M, int n, int64_t *rows) {
  uint32_t c1 = stream_next(s, ST_SAMPLE), w[4];
  int64_t sorted[2 * DREAMZS_MAX_DEPAIRS];
  for (int j = 0; j < n; ++j) {
    if ((j & 3) == 0) stream_block(s, c1, (uint32_t)(j >> 2), w);
    int64_t r = (int64_t)(((uint64_t)w[j & 3] * (uint64_t)(M - j)) >> 32);
    for (int q = 0; q < j; ++q) if (r >= sorted[q]) r += 1;
    rows[j] = r;
    int q = j;
    while (q > 0 && sorted[q - 1] > r) { sorted[q] = sorted[q - 1]; --q; }
    sorted[q] = r;
  }
}
/* np.random.multinomial(1, p) -> index of the 1, by inverse CDF on a running sum */
static int stream_multinomial(stream_t *s, const double *p, int n) {
  double u = stream_uniform53(s, ST_MULTINOMIAL), acc = 0.0;
  for (int j = 0; j < n; ++j) { acc = acc + p[j]; if (u < acc) return j; }
  return n - 1;
}
static int stream_randint(stream_t *s, int n) {
  uint32_t w[4]; stream_block(s, stream_next(s, ST_RANDINT), 0, w);
  return (int)(((uint64_t)w[0] * (uint64_t)n) >> 32);
}

/* ------------------------------------------------------------------ numpy helpers */
/* numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src) so that sums agree
 * with np.sum bit for bit; np.add.reduce seeds the accumulator with a[0]. */
static double pairwise(const double *a, int64_t n) {
  if (n < 8) { double r = 0.0; for (int64_t i = 0; i < n; ++i) r += a[i]; return r; }
  if (n <= 128) {
    double r[8]; int64_t i;
    for (i = 0; i < 8; ++i) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int64_t n2 = n / 2; n2 -= n2 % 8;
  return pairwise(a, n2) + pairwise(a + n2, n - n2);
}
static double np_sum(const double *a, int64_t n) { return n == 0 ? 0.0 : a[0] + pairwise(a + 1, n - 1); }
static double nan_to_num(double x) {
  if (isnan(x)) return 0.0;
  if (isinf(x)) return x > 0 ? DBL_MAX : -DBL_MAX;
  return x;
}
static double dot(const double *a, const double *b, int n) {
  double s = 0.0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s;
}

/* ------------------------------------------------------------------ model: priors + targets */
/* Model.total_logp prior part (pydream/model.py:17-28) with SampledParam.prior
 * (pydream/parameters.py:37-47) for scipy.stats.norm / uniform, FlatParam.prior (:62-63). */
static double log_prior(const dreamzs_config *cfg, const dreamzs_state *st, const double *x, double *tmp) {
  int d = cfg->ndim, any = 0;
  for (int i = 0; i < d; ++i) {
    int k = st->prior_kind ? st->prior_kind[i] : DREAMZS_PRIOR_FLAT;
    double v = 0.0;
    if (k == DREAMZS_PRIOR_NORMAL) {
      double y = (x[i] - st->prior_a[i]) / st->prior_b[i];
      v = (-(y * y) / 2.0 - 0.9189385332046727) - log(st->prior_b[i]);  /* log(sqrt(2 pi)) */
      any = 1;
    } else if (k == DREAMZS_PRIOR_UNIFORM) {
      double y = (x[i] - st->prior_a[i]) / st->prior_b[i];
      v = (y >= 0.0 && y <= 1.0) ? 0.0 - log(st->prior_b[i]) : -INFINITY;
      any = 1;
    }
    tmp[i] = v;
  }
  return any ? np_sum(tmp, d) : 0.0;
}

static double log_like(const dreamzs_config *cfg, const dreamzs_state *st, const double *x, double *tmp) {
  int d = cfg->ndim;
  const double *tb = st->target_table;
  switch (cfg->target_kind) {
    case DREAMZS_TARGET_CONSTANT: return tb[0];
    case DREAMZS_TARGET_SUMSHIFT:  /* pydream/tests/test_models.py:46-50 */
      for (int i = 0; i < d; ++i) tmp[i] = x[i] + tb[0];
      return np_sum(tmp, d);
    case DREAMZS_TARGET_GAUSSIAN_DENSE: {  /* dream_ex_ndim_gaussian.py:49-52 */
      const double *A = tb + 1;
      for (int i = 0; i < d; ++i) tmp[i] = x[i] * dot(A + (size_t)i * d, x, d);
      return tb[0] - .5 * np_sum(tmp, d);
    }
    case DREAMZS_TARGET_MIXTURE: {  /* mixturemodel.py:39-48 */
      double lh[2];
      for (int j = 0; j < 2; ++j) {
        const double *mu = tb + 2 + (size_t)j * d;
        for (int i = 0; i < d; ++i) { double r = x[i] - mu[i]; tmp[i] = r * r; }
        lh[j] = -.5 * np_sum(tmp, d) + tb[j];
      }
      double mx = lh[0] > lh[1] ? lh[0] : lh[1];
      double density = exp(lh[0] - mx) + exp(lh[1] - mx);
      return log(density) + mx;
    }
    case DREAMZS_TARGET_BANANA: {  /* pydream_b200/targets.py Banana (target not from reference) */
      double b = tb[0], v1 = tb[1];
      double y2 = x[1] + b * (x[0] * x[0]) - v1 * b;
      double ss = (x[0] * x[0]) / v1 + y2 * y2;
      if (d > 2) { for (int i = 2; i < d; ++i) tmp[i - 2] = x[i] * x[i]; ss = ss + np_sum(tmp, d - 2); }
      return -.5 * ss;
    }
    default: return NAN;
  }
}

/* ------------------------------------------------------------------ proposals */
typedef struct {
  int run_snooker, cr_idx, delta, lvl_idx;
  double CR;
} decisions_t;

typedef struct {
  double *pts;        /* n x d proposals */
  double *snk_logp;   /* n */
  double *zrow;       /* n x d: snooker z of each point */
  double gamma_any_one; /* 1.0 if any gamma of the batch == 1.0 */
  int64_t rows[DREAMZS_MAX_MULTITRY * 3 > DREAMZS_MAX_MULTITRY * 2 * DREAMZS_MAX_DEPAIRS
                   ? DREAMZS_MAX_MULTITRY * 3 : DREAMZS_MAX_MULTITRY * 2 * DREAMZS_MAX_DEPAIRS];
  int nrows;
} batch_t;

/* Boundary handling, pydream/Dream.py:734-791. */
static void apply_bounds(const dreamzs_config *cfg, const dreamzs_state *st, stream_t *s, double *p, double *tmp) {
  int d = cfg->ndim, nlo = 0, nhi = 0;
  for (int i = 0; i < d; ++i) {
    if (p[i] < st->mins[i]) p[i] = 2 * st->mins[i] - p[i];
    else if (p[i] > st->maxs[i]) p[i] = 2 * st->maxs[i] - p[i];
  }
  /* masks of the second test are both taken before either redraw (Dream.py:769-770) */
  unsigned char *lo = (unsigned char *)(tmp + d), *hi = lo + d;
  for (int i = 0; i < d; ++i) { lo[i] = p[i] < st->mins[i]; hi[i] = p[i] > st->maxs[i]; nlo += lo[i]; nhi += hi[i]; }
  if (nlo) {
    stream_uniform32_vec(s, ST_RAND, nlo, tmp);
    for (int i = 0, j = 0; i < d; ++i) if (lo[i]) p[i] = st->mins[i] + tmp[j++] * (st->maxs[i] - st->mins[i]);
  }
  if (nhi) {
    stream_uniform32_vec(s, ST_RAND, nhi, tmp);
    for (int i = 0, j = 0; i < d; ++i) if (hi[i]) p[i] = st->mins[i] + tmp[j++] * (st->maxs[i] - st->mins[i]);
  }
}

/* generate_proposal_points (pydream/Dream.py:670-796) incl. sample_from_history (:646-668),
 * set_gamma (:601-626) and snooker_update (:798-837). work: >= 8*d doubles. */
static void generate_batch(const dreamzs_config *cfg, const dreamzs_state *st, stream_t *s, int64_t M,
                           const decisions_t *dc, int n, const double *q0, batch_t *out, double *work) {
  const int d = cfg->ndim, ld = cfg->ld;
  const double *Z = st->Z;
  double *zeta = work, *e = work + (size_t)n * d, *U = e + (size_t)n * d, *tmp = U + (size_t)n * d;
  out->gamma_any_one = 0.0;
  out->nrows = 0;
  if (!dc->run_snooker) {
    const int delta = dc->delta;
    for (int p = 0; p < n; ++p) { stream_sample(s, M, 2 * delta, out->rows + out->nrows); out->nrows += 2 * delta; }
    for (int p = 0; p < n; ++p) {
      stream_normal_vec(s, d, zeta + (size_t)p * d);
      for (int i = 0; i < d; ++i) zeta[(size_t)p * d + i] = 0.0 + cfg->zeta * zeta[(size_t)p * d + i];
    }
    for (int p = 0; p < n; ++p) {
      stream_uniform32_vec(s, ST_UNIFORM_VEC, d, e + (size_t)p * d);
      for (int i = 0; i < d; ++i) e[(size_t)p * d + i] = (-cfg->lamb + (cfg->lamb - (-cfg->lamb)) * e[(size_t)p * d + i]) + 1;
    }
    for (int p = 0; p < n; ++p) stream_uniform32_vec(s, ST_UNIFORM_VEC, d, U + (size_t)p * d);
    for (int p = 0; p < n; ++p) {
      const double *Up = U + (size_t)p * d;
      int dprime = 0;
      for (int i = 0; i < d; ++i) dprime += (Up[i] < dc->CR);
      /* set_gamma: the unity draw is always made first (Dream.py:615) */
      double pg[2] = {cfg->p_gamma_unity, 1 - cfg->p_gamma_unity};
      int unity = stream_multinomial(s, pg, 2);
      double gamma;
      if (unity == 0) gamma = 1.0;
      else {
        int di = dprime >= 1 ? dprime - 1 : d - 1;   /* python index -1 -> last entry */
        gamma = st->gamma_table[((size_t)dc->lvl_idx * cfg->nDEpairs + (delta - 1)) * d + di];
      }
      if (gamma == 1.0) out->gamma_any_one = 1.0;
      const int64_t *rw = out->rows + (size_t)p * 2 * delta;
      double *pt = out->pts + (size_t)p * d;
      for (int i = 0; i < d; ++i) {
        double a = Z[(size_t)rw[0] * ld + i], b = Z[(size_t)rw[delta] * ld + i];
        for (int j = 1; j < delta; ++j) { a += Z[(size_t)rw[j] * ld + i]; b += Z[(size_t)rw[delta + j] * ld + i]; }
        double diff = a - b;
        double v = q0[i] + (e[(size_t)p * d + i] * gamma) * diff + zeta[(size_t)p * d + i];
        if (Up[i] > dc->CR) v = q0[i];
        pt[i] = v;
      }
      out->snk_logp[p] = 0.0;
    }
  } else {
    double pg[2] = {cfg->p_gamma_unity, 1 - cfg->p_gamma_unity};
    (void)stream_multinomial(s, pg, 2);
    double gamma = 1.2 + (2.2 - 1.2) * stream_uniform53(s, ST_UNIFORM_SCAL);
    int64_t *zi = out->rows, *zz = out->rows + n;
    for (int p = 0; p < n; ++p) stream_sample(s, M, 1, zi + p);
    for (int p = 0; p < n; ++p) { stream_sample(s, M, 1, zz + 2 * p); stream_sample(s, M, 1, zz + 2 * p + 1); }
    out->nrows = 3 * n;
    for (int p = 0; p < n; ++p) {
      const double *z = Z + (size_t)zi[p] * ld, *z1 = Z + (size_t)zz[2 * p] * ld, *z2 = Z + (size_t)zz[2 * p + 1] * ld;
      double *v = tmp, *t2 = tmp + d, *pt = out->pts + (size_t)p * d;
      for (int i = 0; i < d; ++i) v[i] = q0[i] - z[i];
      double D = dot(v, v, d);
      if (n > 1) {  /* Dream.py:816-822 */
        for (int i = 0; i < d; ++i) t2[i] = (z1[i] - z2[i]) * v[i];
        double sc = np_sum(t2, d) / D;
        for (int i = 0; i < d; ++i) pt[i] = q0[i] + gamma * nan_to_num(sc * v[i]);
      } else {      /* Dream.py:827-833; divide(where=D!=0) pinned to 0 where masked */
        for (int i = 0; i < d; ++i) t2[i] = (D != 0) ? ((z1[i] - z2[i]) * v[i]) / D : 0.0;
        double sc = nan_to_num(np_sum(t2, d));
        for (int i = 0; i < d; ++i) pt[i] = q0[i] + gamma * (sc * v[i]);
      }
      for (int i = 0; i < d; ++i) t2[i] = pt[i] - z[i];
      double norm = sqrt(dot(t2, t2, d));
      out->snk_logp[p] = (norm != 0 ? log(norm) : 0.0) * (d - 1);   /* log(where=False) pinned to 0 */
      memcpy(out->zrow + (size_t)p * d, z, sizeof(double) * d);
    }
    if (gamma == 1.0) out->gamma_any_one = 1.0;
  }
  if (cfg->hardboundaries && st->mins && st->maxs)
    for (int p = 0; p < n; ++p) apply_bounds(cfg, st, s, out->pts + (size_t)p * d, tmp);
}

/* ------------------------------------------------------------------ one chain-step */
typedef struct { double *buf; } scratch_t;

/* Dream.astep (pydream/Dream.py:193-362) for one chain; appends and adaptation are applied by
 * the caller after the sweep.  Returns the decision word (include/dreamzs.h). */
static uint32_t chain_step(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int c_local,
                           int64_t M, const double *cr_probs, const double *gamma_probs, double *q_new,
                           double *work, int64_t *rows_dbg, int rows_dbg_n, double T) {
  const int d = cfg->ndim, ld = cfg->ld, k = cfg->multitry;
  const int c_global = cfg->chain_begin + c_local;
  const double *q0 = st->X + (size_t)c_local * ld;
  stream_t s; stream_init(&s, cfg->seed, (uint32_t)c_global, (uint32_t)iter);
  decisions_t dc; dc.run_snooker = 0;
  if (cfg->snooker != 0) { double ps[2] = {cfg->snooker, 1 - cfg->snooker}; dc.run_snooker = stream_multinomial(&s, ps, 2) == 0; }
  dc.cr_idx = stream_multinomial(&s, cr_probs, cfg->nCR);
  dc.CR = (double)(dc.cr_idx + 1) / (double)cfg->nCR;
  dc.delta = cfg->nDEpairs > 1 ? 1 + stream_randint(&s, cfg->nDEpairs) : 1;
  dc.lvl_idx = stream_multinomial(&s, gamma_probs, cfg->ngamma);

  /* carve work space */
  double *pts = work;                      work += (size_t)k * d;
  double *zrow = work;                     work += (size_t)k * d;
  double *rpts = work;                     work += (size_t)k * d;
  double *rz = work;                       work += (size_t)k * d;
  double *tmp = work;                      work += (size_t)2 * d + 16;
  double *gen = work;                      /* >= (3k+4) d */
  double snk[DREAMZS_MAX_MULTITRY], rsnk[DREAMZS_MAX_MULTITRY];
  batch_t b; b.pts = pts; b.snk_logp = snk; b.zrow = zrow;
  double last_prior = st->last_prior[c_local], last_like = st->last_like[c_local];
  double last_logp = T * last_like + last_prior;   /* Dream.py:243, 268 */
  int accepted = 0, sel = 0, gamma_one;
  double new_prior = 0, new_like = 0;
  const double *q_prop;

  generate_batch(cfg, st, &s, M, &dc, k, q0, &b, gen);
  int dbg_off = 0;
#define DBG_ROWS(bt) do { if (rows_dbg) for (int j = 0; j < (bt).nrows && dbg_off < rows_dbg_n; ++j) rows_dbg[dbg_off++] = (bt).rows[j]; } while (0)
  if (rows_dbg) for (int j = 0; j < rows_dbg_n; ++j) rows_dbg[j] = -1;
  DBG_ROWS(b);
  gamma_one = b.gamma_any_one == 1.0;
  if (k == 1) {
    double q_prior = log_prior(cfg, st, pts, tmp), q_like = log_like(cfg, st, pts, tmp);
    double q_logp = T * q_like + q_prior, mr;   /* Dream.py:274 */
    if (dc.run_snooker) {   /* Dream.py:326-332 */
      for (int i = 0; i < d; ++i) tmp[i] = q0[i] - zrow[i];
      double norm = sqrt(dot(tmp, tmp, d));
      double cur = (norm != 0 ? log(norm) : 0.0) * (d - 1);
      mr = nan_to_num((q_logp + snk[0]) - (last_logp + cur));
    } else mr = nan_to_num(q_logp) - nan_to_num(last_logp);   /* Dream.py:334 */
    if (isfinite(mr)) accepted = log(stream_uniform53(&s, ST_UNIFORM_SCAL)) < mr;  /* metrop_select, Dream.py:980-998 */
    q_prop = pts; new_prior = q_prior; new_like = q_like;
  } else {
    double pri[DREAMZS_MAX_MULTITRY], lik[DREAMZS_MAX_MULTITRY], lps[DREAMZS_MAX_MULTITRY];
    double rpri[DREAMZS_MAX_MULTITRY], rlik[DREAMZS_MAX_MULTITRY], rlps[DREAMZS_MAX_MULTITRY];
    for (int guard = 0;; ++guard) {  /* Dream.py:278-289 */
      int anyfinite = 0;
      for (int p = 0; p < k; ++p) {
        pri[p] = log_prior(cfg, st, pts + (size_t)p * d, tmp);
        lik[p] = log_like(cfg, st, pts + (size_t)p * d, tmp);
        lps[p] = T * lik[p] + pri[p];   /* Dream.py:279, 899 */
        anyfinite |= isfinite(lps[p]) != 0;
      }
      if (anyfinite || guard >= 1000) break;
      generate_batch(cfg, st, &s, M, &dc, k, q0, &b, gen);
      DBG_ROWS(b);
    }
    /* mt_choose_proposal_pt, Dream.py:883-917 */
    double mx = lps[0], w[DREAMZS_MAX_MULTITRY] = {0}, prob[DREAMZS_MAX_MULTITRY], sum = 0.0;
    for (int p = 1; p < k; ++p) if (lps[p] > mx) mx = lps[p];
    for (int p = 0; p < k; ++p) w[p] = exp(lps[p] - mx);
    sum = np_sum(w, k);
    for (int p = 0; p < k; ++p) prob[p] = w[p] / sum;
    sel = stream_multinomial(&s, prob, k);
    q_prop = pts + (size_t)sel * d;
    /* reference points around the selected proposal, Dream.py:295-303 */
    batch_t rb; rb.pts = rpts; rb.snk_logp = rsnk; rb.zrow = rz;
    generate_batch(cfg, st, &s, M, &dc, k - 1, q_prop, &rb, gen);
    gamma_one = rb.gamma_any_one == 1.0;
    DBG_ROWS(rb);
    for (int p = 0; p < k - 1; ++p) {
      rpri[p] = log_prior(cfg, st, rpts + (size_t)p * d, tmp);
      rlik[p] = log_like(cfg, st, rpts + (size_t)p * d, tmp);
    }
    rlik[k - 1] = last_like; rpri[k - 1] = last_prior;   /* Dream.py:877-879 */
    for (int p = 0; p < k; ++p) rlps[p] = T * rlik[p] + rpri[p];   /* Dream.py:303 */
    double tp[DREAMZS_MAX_MULTITRY], trf[DREAMZS_MAX_MULTITRY];
    if (dc.run_snooker) {   /* Dream.py:306-313 */
      rsnk[k - 1] = 0.0;
      for (int p = 0; p < k; ++p) { tp[p] = lps[p] + snk[p]; trf[p] = rlps[p] + rsnk[p] + snk[p]; }
    } else for (int p = 0; p < k; ++p) { tp[p] = lps[p]; trf[p] = rlps[p]; }
    double m2 = tp[0];
    for (int p = 0; p < k; ++p) { if (tp[p] > m2) m2 = tp[p]; if (trf[p] > m2) m2 = trf[p]; }
    double wp[DREAMZS_MAX_MULTITRY], wr[DREAMZS_MAX_MULTITRY];
    for (int p = 0; p < k; ++p) { wp[p] = exp(tp[p] - m2); wr[p] = exp(trf[p] - m2); }
    double mr = nan_to_num(log(np_sum(wp, k) / np_sum(wr, k)));   /* Dream.py:320-323 */
    if (isfinite(mr)) accepted = log(stream_uniform53(&s, ST_UNIFORM_SCAL)) < mr;
    new_prior = pri[sel]; new_like = lik[sel];
  }
  int changed = 0;
  if (accepted) for (int i = 0; i < d; ++i) changed |= (q_prop[i] != q0[i]);
  for (int i = 0; i < d; ++i) q_new[i] = accepted ? q_prop[i] : q0[i];
  if (changed) { st->last_prior[c_local] = new_prior; st->last_like[c_local] = new_like; }  /* Dream.py:336-347 */
  return (uint32_t)changed | ((uint32_t)dc.run_snooker << 1) | ((uint32_t)dc.cr_idx << 2) | ((uint32_t)dc.lvl_idx << 6) |
         ((uint32_t)dc.delta << 10) | ((uint32_t)sel << 14) | ((uint32_t)gamma_one << 18) | ((uint32_t)accepted << 19);
}

int64_t dreamzs_oracle_work_doubles(const dreamzs_config *cfg) {
  return (int64_t)(7 * cfg->multitry + 8) * cfg->ndim + 64;
}

/* First-call branch of astep (Dream.py:266-268). */
int dreamzs_oracle_init_logp(const dreamzs_config *cfg, const dreamzs_state *st) {
  double *tmp = (double *)malloc(sizeof(double) * (2 * cfg->ndim + 16));
  for (int c = 0; c < cfg->nchains_local; ++c) {
    const double *x = st->X + (size_t)c * cfg->ld;
    st->last_prior[c] = log_prior(cfg, st, x, tmp);
    st->last_like[c] = log_like(cfg, st, x, tmp);
  }
  free(tmp);
  return 0;
}

/* Adaptation state shared by all chains (Dream_shared_vars, pydream/core.py:281-297). */
typedef struct dreamzs_oracle_adapt {
  int32_t adapt_crossover, adapt_gamma;
  int64_t crossover_burnin;
  double *cr_probs, *ncr_updates, *delta_m;          /* nCR */
  double *gamma_probs, *ngamma_updates, *delta_m_gamma; /* ngamma */
} dreamzs_oracle_adapt;

/* estimate_crossover_probabilities / estimate_gamma_level_probs for one sweep, chain order
 * (pydream/Dream.py:451-540); current_positions holds chain c in row N-1-c (Dream.py:198-200). */
static void adapt_sweep(const dreamzs_config *cfg, dreamzs_oracle_adapt *ad, int64_t iter, const double *Xold,
                        const double *Xnew, const uint32_t *dec) {
  const int N = cfg->nchains_local, d = cfg->ndim, ld = cfg->ld;
  const int in_window = iter > 10 && iter < ad->crossover_burnin, final = iter == ad->crossover_burnin;
  if (!(in_window || final) || !(ad->adapt_crossover || ad->adapt_gamma)) return;
  double *mean = (double *)calloc(3 * (size_t)d, sizeof(double)), *sd = mean + d, *tmp = sd + d;
  for (int r = 0; r < N; ++r) { const double *x = Xnew + (size_t)(N - 1 - r) * ld; for (int i = 0; i < d; ++i) mean[i] = r ? mean[i] + x[i] : x[i]; }
  for (int i = 0; i < d; ++i) mean[i] /= N;
  for (int r = 0; r < N; ++r) {
    const double *x = Xnew + (size_t)(N - 1 - r) * ld;
    for (int i = 0; i < d; ++i) { double t = x[i] - mean[i]; t = t * t; sd[i] = r ? sd[i] + t : t; }
  }
  for (int i = 0; i < d; ++i) sd[i] = sqrt(sd[i] / N);
  for (int c = 0; c < N; ++c) {
    const uint32_t w = dec[c];
    const int snk = (w >> 1) & 1, gone = (w >> 18) & 1;
    const double *a = Xold + (size_t)c * ld, *b = Xnew + (size_t)c * ld;
    if (ad->adapt_crossover && (final || !gone)) {
      const int m = snk ? cfg->nCR - 1 : (int)((w >> 2) & 15);
      ad->ncr_updates[m] += 1;
      for (int i = 0; i < d; ++i) { double s = sd[i] == 0 ? 1e-12 : sd[i]; double t = (b[i] - a[i]) / s; tmp[i] = t * t; }
      ad->delta_m[m] = ad->delta_m[m] + nan_to_num(np_sum(tmp, d));
      int all = 1; for (int j = 0; j < cfg->nCR; ++j) all &= ad->delta_m[j] != 0;
      if (all) {
        double sum;
        for (int j = 0; j < cfg->nCR; ++j) ad->cr_probs[j] = (ad->delta_m[j] / ad->ncr_updates[j]) * N;
        sum = np_sum(ad->cr_probs, cfg->nCR);
        for (int j = 0; j < cfg->nCR; ++j) ad->cr_probs[j] /= sum;
      }
    }
    if (ad->adapt_gamma && (final || (!gone && !snk))) {
      const int m = (int)((w >> 6) & 15);
      ad->ngamma_updates[m] += 1;
      for (int i = 0; i < d; ++i) { double t = (b[i] - a[i]) / sd[i]; tmp[i] = t * t; }
      ad->delta_m_gamma[m] = ad->delta_m_gamma[m] + nan_to_num(np_sum(tmp, d));
      int all = 1; for (int j = 0; j < cfg->ngamma; ++j) all &= ad->delta_m_gamma[j] != 0;
      if (all) {
        double sum;
        for (int j = 0; j < cfg->ngamma; ++j) ad->gamma_probs[j] = (ad->delta_m_gamma[j] / ad->ngamma_updates[j]) * N;
        sum = np_sum(ad->gamma_probs, cfg->ngamma);
        for (int j = 0; j < cfg->ngamma; ++j) ad->gamma_probs[j] /= sum;
      }
    }
  }
  free(mean);
}

/* worker pool: chains of one sweep are independent (appends/adaptation are deferred), so the
 * sweep is split statically over `nthreads` pthreads with two barriers per iteration. */
typedef struct {
  const dreamzs_config *cfg; const dreamzs_state *st;
  int nthreads; int64_t niter, iter, it, M; int stop;
  double *work, *Xnew; uint32_t *dec; int64_t wd; int64_t *rows_dbg; int rows_dbg_n;
  double crp[DREAMZS_MAX_NCR], gp[DREAMZS_MAX_NGAMMA];
  const double *temperature;   /* per chain; NULL = 1 (no tempering) */
  int64_t rows_per_iter;       /* trace rows per iteration: 1, or 2 under parallel tempering */
  pthread_barrier_t go, done;
} pool_t;
typedef struct { pool_t *pl; int tid; } worker_arg_t;

static void sweep_slice(pool_t *pl, int tid) {
  const dreamzs_config *cfg = pl->cfg;
  const int N = cfg->nchains_local, ld = cfg->ld;
  const int lo = (int)((int64_t)N * tid / pl->nthreads), hi = (int)((int64_t)N * (tid + 1) / pl->nthreads);
  for (int c = lo; c < hi; ++c) {
    memset(pl->Xnew + (size_t)c * ld, 0, sizeof(double) * ld);
    pl->dec[c] = chain_step(cfg, pl->st, pl->iter, c, pl->M, pl->crp, pl->gp, pl->Xnew + (size_t)c * ld,
                            pl->work + (size_t)tid * pl->wd,
                            pl->rows_dbg ? pl->rows_dbg + ((size_t)c * pl->niter + pl->it) * pl->rows_dbg_n : NULL,
                            pl->rows_dbg_n, pl->temperature ? pl->temperature[c] : 1.0);
  }
}
static void *worker_main(void *p) {
  worker_arg_t *a = (worker_arg_t *)p;
  for (;;) {
    pthread_barrier_wait(&a->pl->go);
    if (a->pl->stop) return NULL;
    sweep_slice(a->pl, a->tid);
    pthread_barrier_wait(&a->pl->done);
  }
}

/* The sampling loop of _sample_dream (pydream/core.py:89-129) for all chains in lock-step.
 * st->Z/X/last_* are HOST pointers here.  trace: N x niter x ld, trace_logp: N x niter,
 * decisions: N x niter (may be NULL), rows_dbg: N x niter x rows_dbg_n (may be NULL).
 * *count is Dream_shared_vars.count (rows appended so far), nseed = nseedchains. */
static int run_impl(const dreamzs_config *cfg, const dreamzs_state *st, dreamzs_oracle_adapt *ad,
                    int64_t iter_begin, int64_t niter, int64_t nseed, int64_t *count, double *trace,
                    double *trace_logp, uint32_t *decisions, int64_t *rows_dbg, int32_t rows_dbg_n,
                    int32_t nthreads, const double *temperature, int64_t *swap_pairs) {
  const int N = cfg->nchains_local, ld = cfg->ld;
  const int64_t rpi = temperature ? 2 : 1, TR = niter * rpi;   /* trace rows per iteration / per chain */
  if (cfg->abi_version != DREAMZS_ABI_VERSION || cfg->multitry < 1 || cfg->multitry > DREAMZS_MAX_MULTITRY ||
      cfg->multitry == 2 || cfg->nDEpairs > DREAMZS_MAX_DEPAIRS || cfg->nCR > DREAMZS_MAX_NCR)
    return DREAMZS_E_BADARG;
  /* a shard (nchains_local < nchains_global) steps its own chains only; adaptation sums would need the other shards */
  if (N != cfg->nchains_global && (ad->adapt_crossover || ad->adapt_gamma)) return DREAMZS_E_UNSUPPORTED;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > N) nthreads = N;
  pool_t pl; memset(&pl, 0, sizeof(pl));
  pl.cfg = cfg; pl.st = st; pl.nthreads = nthreads; pl.niter = niter; pl.rows_dbg = rows_dbg; pl.rows_dbg_n = rows_dbg_n;
  pl.temperature = temperature; pl.rows_per_iter = rpi;
  pl.wd = dreamzs_oracle_work_doubles(cfg);
  pl.work = (double *)malloc(sizeof(double) * (size_t)pl.wd * nthreads);
  pl.Xnew = (double *)malloc(sizeof(double) * (size_t)N * ld);
  pl.dec = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)N);
  pthread_t *th = NULL; worker_arg_t *args = NULL;
  if (nthreads > 1) {
    pthread_barrier_init(&pl.go, NULL, (unsigned)nthreads);
    pthread_barrier_init(&pl.done, NULL, (unsigned)nthreads);
    th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    args = (worker_arg_t *)malloc(sizeof(worker_arg_t) * nthreads);
    for (int t = 1; t < nthreads; ++t) { args[t].pl = &pl; args[t].tid = t; pthread_create(&th[t], NULL, worker_main, &args[t]); }
  }
  int rc = DREAMZS_OK;
  for (int64_t it = 0; it < niter && rc == DREAMZS_OK; ++it) {
    const int64_t iter = iter_begin + it, M = nseed + *count;
    const int appends = iter % cfg->history_thin == 0;
    if ((size_t)(M + (appends ? cfg->nchains_global : 0)) > (size_t)st->Z_capacity_rows) { rc = DREAMZS_E_BADARG; break; }
    memcpy(pl.crp, ad->cr_probs, sizeof(double) * cfg->nCR);
    memcpy(pl.gp, ad->gamma_probs, sizeof(double) * cfg->ngamma);
    pl.iter = iter; pl.it = it; pl.M = M;
    if (nthreads > 1) pthread_barrier_wait(&pl.go);
    sweep_slice(&pl, 0);
    if (nthreads > 1) pthread_barrier_wait(&pl.done);
    adapt_sweep(cfg, ad, iter, st->X, pl.Xnew, pl.dec);
    for (int c = 0; c < N; ++c) {
      const double Tc = temperature ? temperature[c] : 1.0;
      memcpy(st->X + (size_t)c * ld, pl.Xnew + (size_t)c * ld, sizeof(double) * ld);
      memcpy(trace + ((size_t)c * TR + it * rpi) * ld, pl.Xnew + (size_t)c * ld, sizeof(double) * ld);
      trace_logp[(size_t)c * TR + it * rpi] = Tc * st->last_like[c] + st->last_prior[c];   /* core.py:115; :178 under tempering */
      if (decisions) decisions[(size_t)c * TR + it * rpi] = pl.dec[c];
    }
    if (appends) {   /* record_history, Dream.py:360-362, 919-938: chain c's row is M + (global chain id) */
      for (int c = 0; c < N; ++c)
        memcpy(st->Z + (size_t)(M + cfg->chain_begin + c) * ld, pl.Xnew + (size_t)c * ld, sizeof(double) * ld);
      *count += cfg->nchains_global;   /* a shard's caller fetches the other shards' rows before the next sweep */
    }
    if (temperature) {
      /* temperature swap of _sample_dream_pt (pydream/core.py:183-218): the parent process draws a pair of
       * chains and a uniform from the stream of the pseudo-chain 0xFFFFFFFF (oracle/philox.py), exchanges
       * state, log-likelihood and log-prior when log(u) < alpha, and records every chain a second time. */
      stream_t ds; stream_init(&ds, cfg->seed, 0xFFFFFFFFu, (uint32_t)iter);
      int64_t pr[2]; stream_sample(&ds, cfg->nchains_global, 2, pr);
      const int a = (int)pr[0], b = (int)pr[1];
      const double T1 = temperature[a], T2 = temperature[b], l1 = st->last_like[a], l2 = st->last_like[b];
      const double alpha = ((T1 * l2) + (T2 * l1)) - ((T1 * l1) + (T2 * l2));
      const int swap = log(stream_uniform53(&ds, ST_UNIFORM_SCAL)) < alpha;
      if (swap_pairs) { swap_pairs[3 * it] = a; swap_pairs[3 * it + 1] = b; swap_pairs[3 * it + 2] = swap; }
      for (int c = 0; c < N; ++c)   /* logpnews, with the temperature of the chain that produced it (core.py:176, 207-208) */
        trace_logp[(size_t)c * TR + it * rpi + 1] = trace_logp[(size_t)c * TR + it * rpi];
      if (swap) {
        for (int i = 0; i < ld; ++i) { double t = st->X[(size_t)a * ld + i]; st->X[(size_t)a * ld + i] = st->X[(size_t)b * ld + i]; st->X[(size_t)b * ld + i] = t; }
        double t = st->last_like[a]; st->last_like[a] = st->last_like[b]; st->last_like[b] = t;
        t = st->last_prior[a]; st->last_prior[a] = st->last_prior[b]; st->last_prior[b] = t;
        t = trace_logp[(size_t)a * TR + it * rpi + 1];
        trace_logp[(size_t)a * TR + it * rpi + 1] = trace_logp[(size_t)b * TR + it * rpi + 1];
        trace_logp[(size_t)b * TR + it * rpi + 1] = t;
      }
      for (int c = 0; c < N; ++c) {
        memcpy(trace + ((size_t)c * TR + it * rpi + 1) * ld, st->X + (size_t)c * ld, sizeof(double) * ld);
        if (decisions) decisions[(size_t)c * TR + it * rpi + 1] = (swap && (c == a || c == b)) ? DREAMZS_DECISION_SWAPPED : 0u;
      }
    }
  }
  if (nthreads > 1) {
    pl.stop = 1;
    pthread_barrier_wait(&pl.go);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&pl.go); pthread_barrier_destroy(&pl.done);
    free(th); free(args);
  }
  free(pl.work); free(pl.Xnew); free(pl.dec);
  return rc;
}

int dreamzs_oracle_run(const dreamzs_config *cfg, const dreamzs_state *st, dreamzs_oracle_adapt *ad,
                       int64_t iter_begin, int64_t niter, int64_t nseed, int64_t *count, double *trace,
                       double *trace_logp, uint32_t *decisions, int64_t *rows_dbg, int32_t rows_dbg_n,
                       int32_t nthreads) {
  return run_impl(cfg, st, ad, iter_begin, niter, nseed, count, trace, trace_logp, decisions, rows_dbg, rows_dbg_n,
                  nthreads, NULL, NULL);
}

/* The loop of _sample_dream_pt (pydream/core.py:131-236): every iteration is one astep per chain at the
 * chain's temperature (recorded), then one proposed temperature swap (recorded again): trace is
 * N x 2 niter x ld, trace_logp / decisions N x 2 niter, log_ps = T like + prior (core.py:176).
 * swap_pairs (optional): niter x 3 = (first chain, second chain, accepted). */
int dreamzs_oracle_run_pt(const dreamzs_config *cfg, const dreamzs_state *st, dreamzs_oracle_adapt *ad,
                          const double *temperature, int64_t iter_begin, int64_t niter, int64_t nseed,
                          int64_t *count, double *trace, double *trace_logp, uint32_t *decisions,
                          int64_t *swap_pairs, int32_t nthreads) {
  if (!temperature || cfg->nchains_local != cfg->nchains_global) return DREAMZS_E_BADARG;
  return run_impl(cfg, st, ad, iter_begin, niter, nseed, count, trace, trace_logp, decisions, NULL, 0, nthreads,
                  temperature, swap_pairs);
}

/* Gelman_Rubin, pydream/convergence.py:3-20.  trace: nchains x nsamples x ld. */
int dreamzs_oracle_gelman_rubin(const double *trace, int64_t nchains, int64_t nsamples, int32_t ndim, int64_t ld,
                                double *rhat) {
  const int64_t nb = nsamples / 2, n = nsamples - nb;
  double *mean = (double *)calloc((size_t)nchains * ndim * 2, sizeof(double)), *var = mean + (size_t)nchains * ndim;
  for (int64_t c = 0; c < nchains; ++c)
    for (int i = 0; i < ndim; ++i) {
      double s = 0, v = 0;
      for (int64_t t = nb; t < nsamples; ++t) s += trace[((size_t)c * nsamples + t) * ld + i];
      s /= n;
      for (int64_t t = nb; t < nsamples; ++t) { double r = trace[((size_t)c * nsamples + t) * ld + i] - s; v += r * r; }
      mean[c * ndim + i] = s; var[c * ndim + i] = v / n;
    }
  for (int i = 0; i < ndim; ++i) {
    double W = 0, mm = 0, B = 0;
    for (int64_t c = 0; c < nchains; ++c) { W += var[c * ndim + i]; mm += mean[c * ndim + i]; }
    W /= nchains; mm /= nchains;
    for (int64_t c = 0; c < nchains; ++c) { double r = mean[c * ndim + i] - mm; B += r * r; }
    B /= nchains;
    rhat[i] = sqrt((W * (1 - (1. / nsamples)) + B) / W);
  }
  free(mean);
  return 0;
}

/* gamma table, Dream.__init__ (pydream/Dream.py:173-179); pinned by test_gamma_array
 * (pydream/tests/test_dream.py:68-76). */
void dreamzs_oracle_gamma_table(int32_t ngamma, int32_t nDEpairs, int32_t ndim, double *out) {
  double dec = 1;
  for (int l = 0; l < ngamma; ++l) {
    for (int dl = 1; dl <= nDEpairs; ++dl)
      for (int i = 1; i <= ndim; ++i) out[((size_t)l * nDEpairs + (dl - 1)) * ndim + (i - 1)] = (2.38 / sqrt(2.0 * dl * i)) / dec;
    dec *= 2;
  }
}
