// Dataflow form of the whitened window kernel (dreamzs_wwin_kernel.cuh): the same arithmetic, draws and column layout
// of work, but the CTA's warps are specialised and stream a window's columns through the stages instead of meeting at
// CTA-wide barriers:
//
//   V warps   the draws.  For item n (one window of one group of chains): `pre` = scalar draws, decisions, archive row
//             indices (+ waiting for rows appended inside this launch), crossover masks and d' -- into the set n&1 of
//             the small per-column arrays -- then `fill` = per (column, chunk) task: the column's first task stages the
//             archive rows by TMA, every task draws zeta / e and leaves J, dx (and the float32 normals) in the column's
//             slots, counting itself into the column tile's counter.  pre(n+1) runs right after fill(n), i.e. while
//             the chains of item n are still running.
//   M warps   one per SM sub-partition: tile t of 8 columns is ready when its counter is full; du = dx^T L for all
//             i-tiles in ascending order on DMMA, written in place (a tile's columns are read by this warp only), then
//             the tile is published.
//   C warps   the chains (LPC lanes per chain, states in registers for the whole launch when the CTA owns one group):
//             iteration i of a chain needs the tile that holds column (i, chain); columns are laid out iteration-major so
//             the chains start as soon as the first tile is through and run behind the V / M stream.
// fill(n+1) starts when every chain has finished item n (the column slots are then free); everything else overlaps.
// Flags and counters live in shared memory; a stage publishes with __threadfence_block() + a volatile store and consumers
// spin on volatile loads (checking the abort word, so a timed-out wait for a peer GPU cannot hang the CTA).
#pragma once
#include "dreamzs_wwin_kernel.cuh"

namespace dreamzs {

constexpr int WF_NMW = 4;          // M warps (one per SM sub-partition)

// shared-memory carve-up (byte offsets) of the dataflow kernel
inline WflowLayout wflow_layout(int d, int ld, int TC, int NB, int ngamma) {
  WflowLayout L;
  L.nch = ld / 4; L.nK = ld / 4; L.nI = (ld + 7) / 8;
  L.ntilesL = wwin_ntiles(ld);
  L.ncolmax = TC * NB + TC;                       // a window's columns + the refresh columns (x of every chain)
  L.ntmax = (L.ncolmax + 7) / 8;
  auto up = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t o = 0;
  L.oL = (int32_t)o;     o += up((size_t)L.ntilesL * 32 * 8);
  L.oW = (int32_t)o;     o += up((size_t)L.ncolmax * ld * 8);            // dx / z / x columns -> du
  L.oJ = (int32_t)o;     o += up((size_t)L.ncolmax * ld * 8);            // J = (e*gamma)*diff | snooker: z
  L.oN = (int32_t)o;     o += up((size_t)L.ncolmax * ld * 4);            // float32 normals of zeta (0 where the dimension is reset)
  L.oGam = (int32_t)o;   o += up((size_t)ngamma * d * 8);
  L.oScr = (int32_t)o;   o += up((size_t)9 * L.ncolmax * 8);             // raw scalar draws [kind][column]
  L.oLogu = (int32_t)o;  o += up((size_t)2 * L.ncolmax * 8);             // from here: two sets (item parity)
  L.oGsn = (int32_t)o;   o += up((size_t)2 * L.ncolmax * 8);
  L.oRows = (int32_t)o;  o += up((size_t)2 * L.ncolmax * 24);            // row indices: DE r1, r2 | snooker z, z1, z2
  L.oMeta = (int32_t)o;  o += up((size_t)2 * L.ncolmax * 4);
  L.oDpr = (int32_t)o;   o += up((size_t)2 * L.ncolmax * 4);
  L.oMask = (int32_t)o;  o += up((size_t)2 * L.ncolmax * L.nch);
  L.oMbar = (int32_t)o;  o += up((size_t)(L.ncolmax + 1) * 8);
  L.oUses = (int32_t)o;  o += up((size_t)L.ncolmax);                     // TMA uses of a column slot (mbarrier phase)
  L.oProbs = (int32_t)o; o += 32 * 8;                                    // [0,16) CR, [16,24) gamma level, 24 snooker, 26 unity
  L.oSync = (int32_t)o;  o += up((size_t)(16 + 2 * L.ntmax + 64) * 4);   // abort, pool counter, pre_done; v2cnt[ntmax]; mdone[ntmax]; cfin[<=64]
  L.oPool = (int32_t)o;
  const size_t cap = 227 * 1024;
  size_t slots = o < cap ? (cap - o) / ((size_t)ld * 8) : 0;
  if (slots > (size_t)L.ncolmax) slots = (size_t)L.ncolmax;
  L.npool = (int32_t)slots;
  o += slots * (size_t)ld * 8;
  L.bytes = (int32_t)o;
  L.m_nch = ((1u << 20) + (uint32_t)L.nch - 1u) / (uint32_t)L.nch;
  return L;
}

__device__ __forceinline__ int ld_vol(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void st_vol(int *p, int v) { *reinterpret_cast<volatile int *>(p) = v; }

// one window of one group of chains; every role walks the same sequence
struct WfItem {
  int64_t wt0, M, trace_row0;
  int wn, blk, grp, chain0, nchains, seq, R, ncol;
  bool append, refresh, last_window, valid;
};

__device__ __forceinline__ void wf_fill(WfItem &it, const StepParams &P, int TC, int ngroups) {
  const int64_t thin = P.cfg.history_thin, t_end = P.iter_begin + P.niter;
  it.valid = it.wt0 < t_end;
  if (!it.valid) return;
  const int64_t nxt = ((it.wt0 + thin - 1) / thin) * thin;            // first appending iteration >= wt0
  it.wn = (int)((t_end < nxt + 1 ? t_end : nxt + 1) - it.wt0);
  it.append = (it.wt0 + it.wn - 1) % thin == 0;
  it.refresh = it.wt0 == 0 || ((it.wt0 - 1) % thin == 0 && ((it.wt0 - 1) / thin) % DREAMZS_GAUSS_REFRESH_WINDOWS == 0);
  it.last_window = it.wt0 + it.wn >= t_end;
  it.M = P.archive_rows + (int64_t)it.blk * P.cfg.nchains_global;
  it.trace_row0 = P.tr.trace_offset + (it.wt0 - P.iter_begin);
  it.chain0 = it.grp * TC;
  it.nchains = min(TC, P.cfg.nchains_local - it.chain0);
  it.R = it.refresh ? it.nchains : 0;
  it.ncol = it.nchains * it.wn;
}
__device__ __forceinline__ void wf_first(WfItem &it, const StepParams &P, int TC, int ngroups) {
  it.wt0 = P.iter_begin; it.blk = 0; it.grp = blockIdx.x; it.seq = 0;
  wf_fill(it, P, TC, ngroups);
}
__device__ __forceinline__ void wf_next(WfItem &it, const StepParams &P, int TC, int ngroups) {
  it.seq += 1;
  it.grp += gridDim.x;
  if (it.grp >= ngroups) {
    it.grp = blockIdx.x;
    if (it.append) it.blk += 1;
    it.wt0 += it.wn;
  }
  wf_fill(it, P, TC, ngroups);
}

template <int LPC>
__global__ void __launch_bounds__(WW_THREADS, 1) dreamzs_wflow_kernel(const StepParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = P.cfg.ndim, ld = P.cfg.ld, TC = P.ww_tc;
  const WflowLayout &L = P.wf_L;
  const int nch = L.nch, nK = L.nK, NCM = L.ncolmax;
  double *Lf = reinterpret_cast<double *>(smem_raw + L.oL), *Wc = reinterpret_cast<double *>(smem_raw + L.oW);
  double *Jc = reinterpret_cast<double *>(smem_raw + L.oJ);
  float *Nz = reinterpret_cast<float *>(smem_raw + L.oN);
  double *gam = reinterpret_cast<double *>(smem_raw + L.oGam);
  uint2 *scr = reinterpret_cast<uint2 *>(smem_raw + L.oScr);
  double *logu2 = reinterpret_cast<double *>(smem_raw + L.oLogu), *gsn2 = reinterpret_cast<double *>(smem_raw + L.oGsn);
  int64_t *rows2 = reinterpret_cast<int64_t *>(smem_raw + L.oRows);
  uint32_t *meta2 = reinterpret_cast<uint32_t *>(smem_raw + L.oMeta);
  int *dpr2 = reinterpret_cast<int *>(smem_raw + L.oDpr);
  unsigned char *mask2 = smem_raw + L.oMask;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + L.oMbar);   // [0, NCM) column slots, [NCM] the factor
  unsigned char *uses = smem_raw + L.oUses;
  double *probs = reinterpret_cast<double *>(smem_raw + L.oProbs);
  int *sync = reinterpret_cast<int *>(smem_raw + L.oSync);
  int *abort_s = sync, *pool_n = sync + 1, *known = sync + 2;         // known[1 + DREAMZS_MAX_PEERS]: append blocks known complete
  int *v2cnt = sync + 16, *mdone = v2cnt + L.ntmax, *cfin = mdone + L.ntmax;
  double *pool = reinterpret_cast<double *>(smem_raw + L.oPool);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int CPW = 32 / LPC;
  const int NCW = (TC + CPW - 1) / CPW;                  // C warps [0, NCW), M warps [NCW, NCW + WF_NMW), V warps the rest
  const int vwarp0 = NCW + WF_NMW, nV = (WW_WARPS - vwarp0) * 32;
  const double logF = P.st.target_table[0];
  const uint32_t row_bytes = (uint32_t)ld * 8u;
  const uint32_t s0 = P.cfg.snooker != 0 ? 1u : 0u;   // multinomial call number of the CR draw
  const uint32_t k0 = (uint32_t)P.cfg.seed, k1 = (uint32_t)(P.cfg.seed >> 32);
  const int ngroups = (P.cfg.nchains_local + TC - 1) / TC;
  const bool resident = ngroups <= (int)gridDim.x;                  // one group per CTA: chain states stay in registers
  volatile int32_t *status = reinterpret_cast<volatile int32_t *>(P.ww_sync);     // word 0: != 0 aborts the launch
  uint32_t *counters = P.ww_sync ? P.ww_sync + 16 : nullptr;                     // chains that have made append #j of this launch
  int dbg_n = 0;
#define WF_STAMP(base) do { if (P.dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 20) P.dbg[(base) + dbg_n] = clock64(); ++dbg_n; } while (0)

  // ---- prologue (all warps): tables, sync words, mbarriers; one TMA bulk copy brings the packed factor
  if (tid <= NCM) mbar_init(mbar + tid, 1);
  if (tid < 32) {
    double v = 0.0;
    if (tid < 16) v = tid < P.cfg.nCR ? P.st.cr_probs[tid] : 0.0;
    else if (tid < 24) v = tid - 16 < P.cfg.ngamma ? P.st.gamma_probs[tid - 16] : 0.0;
    else if (tid == 24) v = P.cfg.snooker;
    else if (tid == 26) v = P.cfg.p_gamma_unity;
    probs[tid] = v;
  }
  for (int i = tid; i < 16 + 2 * L.ntmax + 64; i += WW_THREADS) sync[i] = 0;
  for (int i = tid; i < NCM; i += WW_THREADS) uses[i] = 0;
  for (int i = tid; i < P.cfg.ngamma * d; i += WW_THREADS) {   // gamma_table[level][0][:] (one DE pair)
    const int lv = i / d;
    gam[i] = P.st.gamma_table[(size_t)lv * P.cfg.nDEpairs * d + (i - lv * d)];
  }
  __syncthreads();
  if (tid == 32) {
    // a peer's append has not arrived within the timeout (in an earlier launch, or now): leave everything untouched,
    // the host raises (DreamEngine.check_peers)
    bool bad = (P.peer_error && *reinterpret_cast<volatile int32_t *>(P.peer_error) != 0) || (status && *status != 0);
    if (!bad && P.wait_k) bad = !peer_wait(P.my_flags, P.world, P.my_rank, P.wait_k, P.peer_error);   // the peers' rows have landed
    if (bad) *abort_s = 1;
  }
  __syncthreads();
  if (ld_vol(abort_s)) return;
  if (tid == 0) {
    fence_proxy_async();
    const uint32_t bytes = (uint32_t)L.ntilesL * 256u;
    mbar_expect_tx(mbar + NCM, bytes);
    tma_load_row(Lf, P.st.gauss_L, bytes, mbar + NCM);
  }

  WfItem it;
  wf_first(it, P, TC, ngroups);

  if (warp >= vwarp0) {
    // ======================================================================================== V warps: the draws
    const int vtid = tid - vwarp0 * 32;
    const RowWait rw = {P.archive_rows, P.cfg.nchains_global, P.cfg.chain_begin, P.cfg.nchains_local, P.ww_k0, P.my_flags,
                        P.peer_error, counters, status, nullptr};
    // pre(item): scalar draws -> decisions, row indices, masks, d' in set item.seq & 1
    auto pre = [&](const WfItem &w) {
      const int set = w.seq & 1;
      double *logu = logu2 + set * NCM, *gsn = gsn2 + set * NCM;
      int64_t *rows = rows2 + (size_t)set * NCM * 3;
      uint32_t *meta = meta2 + set * NCM;
      int *dpr = dpr2 + set * NCM;
      unsigned char *maskb = mask2 + (size_t)set * NCM * nch;
      const int ncol = w.ncol, nchn = w.nchains;
      const uint32_t m_nc = fdiv20_magic(nchn);
      // ---- S: one Philox block per (kind, column); column c = iteration * nchains + chain (iteration-major)
      for (int task = vtid; task < 9 * ncol; task += nV) {
        int kind = 0, col = task;
        while (col >= ncol) { col -= ncol; ++kind; }
        const int itb = fdiv20(col, m_nc), ch = col - itb * nchn;
        const uint32_t iter = (uint32_t)(w.wt0 + itb);
        const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + w.chain0 + ch);
        uint32_t call = 0, st = ST_MULTINOMIAL;
        const double *pp = probs + 24;
        int n = 2;
        if (kind == 1) { call = s0; pp = probs; n = P.cfg.nCR; }
        else if (kind == 2) { call = s0 + 1; pp = probs + 16; n = P.cfg.ngamma; }
        else if (kind == 3) { call = s0 + 2; pp = probs + 26; }
        else if (kind == 4) { st = ST_UNIFORM_SCAL; }
        else if (kind == 5) { call = 1; st = ST_UNIFORM_SCAL; }
        else if (kind >= 6) { call = (uint32_t)(kind - 6); st = ST_SAMPLE; }
        const uint4 wd = philox_inl(0u, (call << 3) | st, iter, c_global, k0, k1);
        uint2 out = make_uint2(wd.x, wd.y);
        if (kind < 4) {          // np.random.multinomial(1, p): inverse CDF on a running sum
          const double u = u53_of(wd.x, wd.y);
          double acc = 0.0;
          int idx = n - 1;
          bool found = false;
          for (int j = 0; j < n; ++j) {
            acc = acc + pp[j];
            if (!found && u < acc) { idx = j; found = true; }
          }
          out.x = (uint32_t)idx;
        }
        if (kind == 4 || kind == 5) {   // both candidates for the Metropolis uniform: log u now, off the per-column path
          const double lg = log(u53_of(wd.x, wd.y));
          if (kind == 5) out = make_uint2((uint32_t)__double2loint(lg), (uint32_t)__double2hiint(lg));
          else logu[col] = lg;
        }
        scr[kind * NCM + col] = out;
      }
      if (vtid == 0) *pool_n = 0;
      named_sync(1, nV);
      // ---- one thread per column: decisions, archive rows (waiting for rows appended inside this launch)
      for (int col = vtid; col < ncol; col += nV) {
        const uint2 q0 = scr[col], q1 = scr[NCM + col], q2 = scr[2 * NCM + col], q3 = scr[3 * NCM + col];
        const uint2 u4 = scr[4 * NCM + col], u5 = scr[5 * NCM + col];
        const uint2 r6 = scr[6 * NCM + col], r7 = scr[7 * NCM + col], r8 = scr[8 * NCM + col];
        const bool snk = (s0 != 0u) && q0.x == 0u;
        const int slot = w.R + col;
        const uint32_t use = uses[slot];
        uses[slot] = (unsigned char)(use + 1u);
        // meta word: bits 0-3 CR index, 4-7 gamma level, 8 snooker, 9 gamma == 1 (set by fill), 10 "not unity",
        // 11 parity of the mbarrier phase this use of the slot completes
        uint32_t mt = q1.x | (q2.x << 4) | (snk ? 256u : 0u) | (q3.x != 0u ? 1024u : 0u) | ((use & 1u) << 11);
        if (snk) logu[col] = __hiloint2double((int)u5.y, (int)u5.x);   // 2nd np.random.uniform() after a snooker gamma
        bool ok;
        if (!snk) {
          const int64_t ra = (int64_t)(((uint64_t)r6.x * (uint64_t)w.M) >> 32);
          int64_t rb = (int64_t)(((uint64_t)r6.y * (uint64_t)(w.M - 1)) >> 32);
          if (rb >= ra) rb += 1;
          rows[3 * col] = ra; rows[3 * col + 1] = rb;
          ok = rows_ready(rw, known, ra, rb, -1);
          dpr[col] = 0;
        } else {
          const double g = 1.2 + (2.2 - 1.2) * u53_of(u4.x, u4.y);             // snooker gamma, Dream.py:618
          gsn[col] = g;
          if (g == 1.0) mt |= 512u;
          const int64_t rz = (int64_t)(((uint64_t)r6.x * (uint64_t)w.M) >> 32);
          const int64_t r1 = (int64_t)(((uint64_t)r7.x * (uint64_t)w.M) >> 32), r2 = (int64_t)(((uint64_t)r8.x * (uint64_t)w.M) >> 32);
          rows[3 * col] = rz; rows[3 * col + 1] = r1; rows[3 * col + 2] = r2;
          ok = rows_ready(rw, known, rz, r1, r2);
          const int ps = atomicAdd(pool_n, 1);                               // z1 - z2 goes to the pool when a slot is left
          dpr[col] = ps < L.npool ? ps : -1;
        }
        if (!ok) st_vol(abort_s, 1);
        meta[col] = mt;
      }
      named_sync(1, nV);
      // ---- V1: crossover uniforms -> keep mask, d'
      const int ntask = ncol * nch;
      for (int task = vtid; task < ntask; task += nV) {
        const int col = fdiv20(task, L.m_nch), q = task - col * nch;
        const uint32_t mt = meta[col];
        if (mt & 256u) continue;
        const int itb = fdiv20(col, m_nc), ch = col - itb * nchn;
        const uint32_t iter = (uint32_t)(w.wt0 + itb);
        const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + w.chain0 + ch);
        const uint4 wu = philox_inl((uint32_t)q, (1u << 3) | ST_UNIFORM_VEC, iter, c_global, k0, k1);
        // U = w 2^-32 exactly, so U < CR <=> w < ceil(CR 2^32) and U > CR <=> w > floor(CR 2^32)
        const double CRs = ((double)((mt & 15u) + 1u) / (double)P.cfg.nCR) * 4294967296.0;
        const uint64_t t_lt = (uint64_t)ceil(CRs), t_gt = (uint64_t)floor(CRs);
        const uint32_t wv[4] = {wu.x, wu.y, wu.z, wu.w};
        unsigned reset = 0;
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (4 * q + j < d) {
            cnt += ((uint64_t)wv[j] < t_lt);
            if ((uint64_t)wv[j] > t_gt) reset |= 1u << j;
          } else reset |= 1u << j;
        }
        maskb[task] = (unsigned char)reset;
        if (cnt) atomicAdd(dpr + col, cnt);
      }
      named_sync(1, nV);
    };

    pre(it);
    int vdbg = 0;
    while (it.valid) {
      if (ld_vol(abort_s)) return;     // uniform over the V warps: set before a named barrier, read after it
      const int set = it.seq & 1;
      double *gsn = gsn2 + set * NCM;
      (void)gsn;
      const int64_t *rows = rows2 + (size_t)set * NCM * 3;
      uint32_t *meta = meta2 + set * NCM;
      const int *dpr = dpr2 + set * NCM;
      const unsigned char *maskb = mask2 + (size_t)set * NCM * nch;
      const int ncol = it.ncol, nchn = it.nchains;
      const uint32_t m_nc = fdiv20_magic(nchn);
      // the column slots are free once every chain has finished the previous item
      for (int c = lane; c < NCW; c += 32)
        while (ld_vol(cfin + c) < it.seq) { __nanosleep(64); }
      __syncwarp();
      if (P.dbg && blockIdx.x == 0 && vtid == 0 && vdbg < 20) P.dbg[20 + vdbg++] = clock64();
      // ---- fill(item): per (column, chunk) task, columns in slot order so that the tiles complete in order
      const int ntask = ncol * nch;
      for (int task = vtid; task < ntask; task += nV) {
        const int col = fdiv20(task, L.m_nch), q = task - col * nch;
        const int slot = it.R + col;
        const uint32_t mt = meta[col];
        const uint32_t parity = (mt >> 11) & 1u;
        double *js = Jc + (size_t)slot * ld + 4 * q, *ws = Wc + (size_t)slot * ld + 4 * q;
        if (q == 0) {    // the column's first task stages its archive rows: z_r1 -> J slot, z_r2 -> W slot | snooker: z -> both
          fence_proxy_async();   // the slots' earlier generic-proxy accesses and the acquired rows are ordered before the async copies
          mbar_expect_tx(mbar + slot, 2u * row_bytes);
          const int64_t ra = rows[3 * col], rb = (mt & 256u) ? ra : rows[3 * col + 1];
          tma_load_row(js, P.st.Z + (size_t)ra * ld, row_bytes, mbar + slot);
          tma_load_row(ws, P.st.Z + (size_t)rb * ld, row_bytes, mbar + slot);
        }
        if (mt & 256u) {   // snooker column: z1 - z2 (Dream.py:810) -> pool slot; z / L^T z arrive by TMA
          const int ps = dpr[col];
          if (ps >= 0) {
            const double *z1 = P.st.Z + (size_t)rows[3 * col + 1] * ld + 4 * q, *z2 = P.st.Z + (size_t)rows[3 * col + 2] * ld + 4 * q;
            const double2 p01 = __ldg(reinterpret_cast<const double2 *>(z1)), p23 = __ldg(reinterpret_cast<const double2 *>(z1) + 1);
            const double2 q01 = __ldg(reinterpret_cast<const double2 *>(z2)), q23 = __ldg(reinterpret_cast<const double2 *>(z2) + 1);
            double2 *bs = reinterpret_cast<double2 *>(pool + (size_t)ps * ld + 4 * q);
            bs[0] = make_double2(p01.x - q01.x, p01.y - q01.y);
            bs[1] = make_double2(p23.x - q23.x, p23.y - q23.y);
          }
          mbar_wait(mbar + slot, parity);
        } else {
          const int itb = fdiv20(col, m_nc), ch = col - itb * nchn;
          const uint32_t iter = (uint32_t)(it.wt0 + itb);
          const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + it.chain0 + ch);
          // two independent Philox blocks, interleaved by the compiler
          const uint4 wn = philox_inl((uint32_t)q, (0u << 3) | ST_NORMAL, iter, c_global, k0, k1);
          const uint4 we = philox_inl((uint32_t)q, (0u << 3) | ST_UNIFORM_VEC, iter, c_global, k0, k1);
          float nz[4];
          normal4f(wn, nz);
          const uint32_t wev[4] = {we.x, we.y, we.z, we.w};
          const unsigned reset = maskb[task];
          const int dprime = dpr[col];
          double gamma = 1.0;
          if (mt & 1024u) gamma = gam[((mt >> 4) & 15u) * d + (dprime >= 1 ? dprime - 1 : d - 1)];
          if (q == 0 && gamma == 1.0) atomicOr(meta + col, 512u);
          float *ns = Nz + (size_t)slot * ld + 4 * q;
          mbar_wait(mbar + slot, parity);
          const double2 a01 = *reinterpret_cast<const double2 *>(js), a23 = *reinterpret_cast<const double2 *>(js + 2);
          const double2 b01 = *reinterpret_cast<const double2 *>(ws), b23 = *reinterpret_cast<const double2 *>(ws + 2);
          const double diff[4] = {a01.x - b01.x, a01.y - b01.y, a23.x - b23.x, a23.y - b23.y};
          double J[4], dl[4];
          float nk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool keep = !((reset >> j) & 1u);
            const double e = (-P.cfg.lamb + (P.cfg.lamb - (-P.cfg.lamb)) * u32_of(wev[j])) + 1;
            const double zt = 0.0 + P.cfg.zeta * (double)nz[j];
            J[j] = keep ? (e * gamma) * diff[j] : 0.0;
            nk[j] = keep ? __fadd_rn(nz[j], 0.0f) : 0.0f;      // -0 -> +0: the chain then needs no `0.0 +` (np.random.normal(0, zeta))
            dl[j] = keep ? J[j] + zt : 0.0;
          }
          *reinterpret_cast<double2 *>(js) = make_double2(J[0], J[1]); *reinterpret_cast<double2 *>(js + 2) = make_double2(J[2], J[3]);
          *reinterpret_cast<double2 *>(ws) = make_double2(dl[0], dl[1]); *reinterpret_cast<double2 *>(ws + 2) = make_double2(dl[2], dl[3]);
          *reinterpret_cast<float4 *>(ns) = make_float4(nk[0], nk[1], nk[2], nk[3]);
        }
        __threadfence_block();
        atomicAdd(v2cnt + (slot >> 3), 1);      // the task is in: its tile is complete at (columns in the tile) x nch
      }
      if (P.dbg && blockIdx.x == 0 && vtid == 0 && vdbg < 20) P.dbg[20 + vdbg++] = clock64();
      // ---- next item: its pre runs now, while the chains of this one are still going
      wf_next(it, P, TC, ngroups);
      if (it.valid) pre(it);
      if (P.dbg && blockIdx.x == 0 && vtid == 0 && vdbg < 20) P.dbg[20 + vdbg++] = clock64();
    }
    return;
  }

  if (warp >= NCW) {
    // ======================================================================================== M warps: du = dx^T L
    const int m = warp - NCW;
    const int l4 = lane >> 2, lm = lane & 3;
    mbar_wait(mbar + NCM, 0);   // the factor has landed
    for (; it.valid; wf_next(it, P, TC, ngroups)) {
      const int ncols_total = it.R + it.ncol;
      const int NT = (ncols_total + 7) >> 3;
      for (int nt = m; nt < NT; nt += WF_NMW) {
        const int target = min(8, ncols_total - 8 * nt) * nch;
        while (ld_vol(v2cnt + nt) < target) { if (ld_vol(abort_s)) return; __nanosleep(32); }
        __threadfence_block();
        const int crow = min(nt * 8 + l4, ncols_total - 1);     // padding rows alias the last column (results dropped)
        double *arow = Wc + (size_t)crow * ld;
        const bool wr = nt * 8 + l4 < ncols_total;
        // i-tiles in ascending pairs; tile I reads j >= 8 I only, so its results may overwrite [8 I, 8 I + 8) at once
        for (int I0 = 0; I0 < L.nI; I0 += 2) {
          const bool two = I0 + 1 < L.nI;
          const int ka = 2 * I0, kb = two ? min(2 * I0 + 2, nK) : nK;   // k in [ka, kb) feeds I0 only, [kb, nK) feeds both
          const double *b0 = Lf + (size_t)wwin_tile0(nK, I0) * 32 + lane;
          const double *b1 = Lf + (size_t)wwin_tile0(nK, two ? I0 + 1 : I0) * 32 + lane;
          double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0, c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
          for (int k = ka; k < kb; ++k) dmma884(a00, a01, arow[4 * k + lm], b0[(size_t)(k - ka) * 32]);
          if (two) {
            int k = kb;
            for (; k + 1 < nK; k += 2) {      // two accumulator sets: four independent DMMA chains
              const double x0 = arow[4 * k + lm], x1 = arow[4 * k + 4 + lm];
              dmma884(a00, a01, x0, b0[(size_t)(k - ka) * 32]);
              dmma884(a10, a11, x0, b1[(size_t)(k - kb) * 32]);
              dmma884(c00, c01, x1, b0[(size_t)(k + 1 - ka) * 32]);
              dmma884(c10, c11, x1, b1[(size_t)(k + 1 - kb) * 32]);
            }
            if (k < nK) {
              const double x0 = arow[4 * k + lm];
              dmma884(a00, a01, x0, b0[(size_t)(k - ka) * 32]);
              dmma884(a10, a11, x0, b1[(size_t)(k - kb) * 32]);
            }
            a00 += c00; a01 += c01; a10 += c10; a11 += c11;
          }
          __syncwarp();   // every lane has read the j-range before it is overwritten
          if (wr) {
            if (8 * I0 + 2 * lm < ld) *reinterpret_cast<double2 *>(arow + 8 * I0 + 2 * lm) = make_double2(a00, a01);
            if (two && 8 * I0 + 8 + 2 * lm < ld) *reinterpret_cast<double2 *>(arow + 8 * I0 + 8 + 2 * lm) = make_double2(a10, a11);
          }
          __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) { st_vol(v2cnt + nt, 0); st_vol(mdone + nt, it.seq + 1); }
      }
      // the next item's tiles are filled only after every chain has finished this one, i.e. after its tiles were consumed
    }
    return;
  }

  // ========================================================================================== C warps: the chains
  {
    const int sub = lane / LPC, g = lane - sub * LPC;
    const int cs = warp * CPW + sub;                  // chain slot in the group
    const int i0 = 4 * g;
    const bool own = i0 < ld;
    const double zeta = P.cfg.zeta;
    double x0[4] = {0, 0, 0, 0}, u0[4] = {0, 0, 0, 0};
    double last_prior = 0.0, last_like = 0.0, ntn_last = 0.0;
    for (; it.valid; wf_next(it, P, TC, ngroups)) {
      const int set = it.seq & 1;
      const double *logu = logu2 + set * NCM, *gsn = gsn2 + set * NCM;
      const int64_t *rows = rows2 + (size_t)set * NCM * 3;
      const uint32_t *meta = meta2 + set * NCM;
      const int *dpr = dpr2 + set * NCM;
      const int nchn = it.nchains;
      const bool cwarp = warp * CPW < nchn;
      if (cwarp) {
        const bool valid = cs < nchn;
        const int csv = valid ? cs : nchn - 1;          // lanes of a missing chain shadow the last one (no stores)
        const int c_local = it.chain0 + csv;
        const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);
        if (!resident || it.seq == 0) {   // chain state -> registers
          if (own) {
            const double2 *xr = reinterpret_cast<const double2 *>(P.st.X + (size_t)c_local * ld + i0);
            const double2 a = xr[0], b = xr[1];
            x0[0] = a.x; x0[1] = a.y; x0[2] = b.x; x0[3] = b.y;
            if (!it.refresh) {
              const double2 *ur = reinterpret_cast<const double2 *>(P.st.gauss_U + (size_t)c_local * ld + i0);
              const double2 c = ur[0], e = ur[1];
              u0[0] = c.x; u0[1] = c.y; u0[2] = e.x; u0[3] = e.y;
            }
          }
          last_prior = P.st.last_prior[c_local];
          last_like = P.st.last_like[c_local];
          ntn_last = nan_to_num(1.0 * last_like + last_prior);
        }
        if (it.refresh) {
          // u = L^T x: x goes through the product stage as column `csv`.  The slot may still be in use by ANOTHER C warp's
          // chain of the previous item (groups differ in size), so wait -- as the V warps do -- until every C warp is past it.
          for (int c = lane; c < NCW; c += 32)
            while (ld_vol(cfin + c) < it.seq) { if (ld_vol(abort_s)) break; __nanosleep(64); }
          __syncwarp();
          if (ld_vol(abort_s)) return;
          if (own && valid) {
            double2 *wd = reinterpret_cast<double2 *>(Wc + (size_t)csv * ld + i0);
            wd[0] = make_double2(x0[0], x0[1]); wd[1] = make_double2(x0[2], x0[3]);
          }
          __threadfence_block();
          __syncwarp();
          if (g == 0 && valid) atomicAdd(v2cnt + (csv >> 3), nch);
          while (ld_vol(mdone + (csv >> 3)) < it.seq + 1) { if (ld_vol(abort_s)) return; __nanosleep(32); }
          __threadfence_block();
          if (own) {
            const double2 *ud = reinterpret_cast<const double2 *>(Wc + (size_t)csv * ld + i0);
            const double2 c = ud[0], e = ud[1];
            u0[0] = c.x; u0[1] = c.y; u0[2] = e.x; u0[3] = e.y;
          }
        }
        double *trow_ptr = P.tr.trace + ((size_t)c_local * P.tr.trace_iters + it.trace_row0) * ld + i0;
        double *lrow_ptr = P.tr.trace_logp + (size_t)c_local * P.tr.trace_iters + it.trace_row0;
        uint32_t *drow_ptr = P.tr.decisions ? P.tr.decisions + (size_t)c_local * P.tr.trace_iters + it.trace_row0 : nullptr;
#pragma unroll 1
        for (int itb = 0; itb < it.wn; ++itb) {
          const int col = itb * nchn + csv, slot = it.R + col;
          while (ld_vol(mdone + (slot >> 3)) < it.seq + 1) { if (ld_vol(abort_s)) return; __nanosleep(32); }
          __threadfence_block();
          const uint32_t mt = meta[col];
          const double lu = logu[col];
          const int run_snooker = (mt >> 8) & 1;
          const double *js = Jc + (size_t)slot * ld + i0, *ws = Wc + (size_t)slot * ld + i0;
          double prop[4] = {0, 0, 0, 0}, un[4] = {0, 0, 0, 0}, snk_logp = 0.0, cur = 0.0;
          if (own) {
            // prop = q0 + e*gamma*diff + zeta (Dream.py:717); Q(prop) = |u + L^T dx|^2
            // (zeta = 0.0 + zeta * n: the stored normals carry no -0, so the product needs no `0.0 +`)
            const double2 j01 = *reinterpret_cast<const double2 *>(js), j23 = *reinterpret_cast<const double2 *>(js + 2);
            const double2 w01 = *reinterpret_cast<const double2 *>(ws), w23 = *reinterpret_cast<const double2 *>(ws + 2);
            const float4 nn = *reinterpret_cast<const float4 *>(Nz + (size_t)slot * ld + i0);
            prop[0] = (x0[0] + j01.x) + zeta * (double)nn.x;
            prop[1] = (x0[1] + j01.y) + zeta * (double)nn.y;
            prop[2] = (x0[2] + j23.x) + zeta * (double)nn.z;
            prop[3] = (x0[3] + j23.y) + zeta * (double)nn.w;
            un[0] = u0[0] + w01.x; un[1] = u0[1] + w01.y; un[2] = u0[2] + w23.x; un[3] = u0[3] + w23.y;
          }
          double part = 0.0;
          if (__any_sync(0xffffffffu, run_snooker)) {
            // snooker_update, Dream.py:827-835 (single-point form); J slot = z, W slot = L^T z, z1 - z2 from the pool
            double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
            if (run_snooker && own) {
              const double2 j01 = *reinterpret_cast<const double2 *>(js), j23 = *reinterpret_cast<const double2 *>(js + 2);
              a[0] = j01.x; a[1] = j01.y; a[2] = j23.x; a[3] = j23.y;
              const int ps = dpr[col];
              if (ps >= 0) {      // staged by the V warps
                const double2 b01 = *reinterpret_cast<const double2 *>(pool + (size_t)ps * ld + i0);
                const double2 b23 = *reinterpret_cast<const double2 *>(pool + (size_t)ps * ld + i0 + 2);
                b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
              } else {            // pool full: read the two rows here
                const double *z1 = P.st.Z + (size_t)rows[3 * col + 1] * ld + i0, *z2 = P.st.Z + (size_t)rows[3 * col + 2] * ld + i0;
                const double2 p01 = __ldg(reinterpret_cast<const double2 *>(z1)), p23 = __ldg(reinterpret_cast<const double2 *>(z1) + 1);
                const double2 q01 = __ldg(reinterpret_cast<const double2 *>(z2)), q23 = __ldg(reinterpret_cast<const double2 *>(z2) + 1);
                b[0] = p01.x - q01.x; b[1] = p01.y - q01.y; b[2] = p23.x - q23.x; b[3] = p23.y - q23.y;
              }
            }
            const double gamma = gsn[col];
            double v[4];
            double D = 0.0, S = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[j] = (i0 + j < d) ? x0[j] - a[j] : 0.0;
              D = fma(v[j], v[j], D);
              b[j] = b[j] * v[j];
            }
            // |q0 - z|^2 and (z1 - z2).(q0 - z) in ONE pass of shuffles; the projection coefficient is then their quotient
            // (Dream.py:829-831 divides element-wise and sums: same value up to the rounding of the summation order;
            // 0 where D == 0, as the masked divide leaves it)
#pragma unroll
            for (int j = 0; j < 4; ++j) S += b[j];
#pragma unroll
            for (int o = LPC / 2; o > 0; o >>= 1) {
              D += __shfl_xor_sync(0xffffffffu, D, o);
              S += __shfl_xor_sync(0xffffffffu, S, o);
            }
            const double sc = (D != 0) ? nan_to_num(S / D) : 0.0;
            const double cg = gamma * sc;
            double nn = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const bool okd = i0 + j < d;
              const double o = okd ? x0[j] + gamma * (sc * v[j]) : 0.0;
              const double ww = okd ? o - a[j] : 0.0;
              nn = fma(ww, ww, nn);
              if (run_snooker) prop[j] = o;
            }
            if (run_snooker && own) {   // L^T dx = c (u - L^T z)
              const double2 w01 = *reinterpret_cast<const double2 *>(ws), w23 = *reinterpret_cast<const double2 *>(ws + 2);
              un[0] = u0[0] + cg * (u0[0] - w01.x); un[1] = u0[1] + cg * (u0[1] - w01.y);
              un[2] = u0[2] + cg * (u0[2] - w23.x); un[3] = u0[3] + cg * (u0[3] - w23.y);
            }
            part = fma(un[1], un[1], un[0] * un[0]) + fma(un[3], un[3], un[2] * un[2]);
            // |prop - z|^2 and Q' in one pass of shuffles
#pragma unroll
            for (int o = LPC / 2; o > 0; o >>= 1) {
              nn += __shfl_xor_sync(0xffffffffu, nn, o);
              part += __shfl_xor_sync(0xffffffffu, part, o);
            }
            if (run_snooker) {
              const double norm = sqrt(nn);
              snk_logp = (norm != 0 ? log(norm) : 0.0) * (d - 1);
              const double n0 = sqrt(D);
              cur = (n0 != 0 ? log(n0) : 0.0) * (d - 1);
            }
          } else {
            part = fma(un[1], un[1], un[0] * un[0]) + fma(un[3], un[3], un[2] * un[2]);
            part = lsum<LPC>(part);
          }
          const double Qn = part;
          int anydiff = (prop[0] != x0[0]) | (prop[1] != x0[1]) | (prop[2] != x0[2]) | (prop[3] != x0[3]);
          if (LPC == 32) anydiff = __any_sync(0xffffffffu, anydiff);
          else {
            const unsigned bal = __ballot_sync(0xffffffffu, anydiff);
            anydiff = ((bal >> (sub * LPC)) & ((LPC == 32) ? 0xffffffffu : ((1u << LPC) - 1u))) != 0u;
          }
          const double q_like = logF - .5 * Qn;
          // mr = nan_to_num(q_logp) - nan_to_num(last_logp) (Dream.py:334); nan_to_num is the identity on finite values,
          // which one comparison establishes (|x| <= DBL_MAX is false for inf and nan)
          double mr = q_like - ntn_last;
          if (!(fabs(q_like) <= DBL_MAX)) mr = nan_to_num(q_like) - ntn_last;
          if (run_snooker) mr = nan_to_num((q_like + snk_logp) - ((1.0 * last_like + last_prior) + cur));   // Dream.py:326-332
          const bool accepted = (fabs(mr) <= DBL_MAX) && lu < mr;                          // metrop_select, Dream.py:980-998
          const int changed = accepted && anydiff;
          if (changed) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { x0[j] = prop[j]; u0[j] = un[j]; }
            last_prior = 0.0;
            last_like = q_like;
            ntn_last = (fabs(q_like) <= DBL_MAX) ? q_like : nan_to_num(q_like);
          }
          const bool appending = it.append && itb == it.wn - 1;
          if (own && valid) {
            *reinterpret_cast<double2 *>(trow_ptr) = make_double2(x0[0], x0[1]);
            *reinterpret_cast<double2 *>(trow_ptr + 2) = make_double2(x0[2], x0[3]);
            if (appending) {   // record_history: the last iteration of the window
              double *zr = P.st.Z + (size_t)(it.M + c_global) * ld + i0;
              *reinterpret_cast<double2 *>(zr) = make_double2(x0[0], x0[1]);
              *reinterpret_cast<double2 *>(zr + 2) = make_double2(x0[2], x0[3]);
              for (int pz = 0; pz < P.npeers; ++pz) {   // replicas over NVLink
                double *zp = P.peer_Z[pz] + (size_t)(it.M + c_global) * ld + i0;
                *reinterpret_cast<double2 *>(zp) = make_double2(x0[0], x0[1]);
                *reinterpret_cast<double2 *>(zp + 2) = make_double2(x0[2], x0[3]);
              }
            }
          }
          if (appending && (counters || P.publish_k)) {
            // the row is visible (to the GPU; to the peers when there are any) before the chain counts itself
            if (P.npeers) __threadfence_system(); else __threadfence();
            __syncwarp();
            if (g == 0 && valid) {
              if (counters) {
                const uint32_t old = atomicAdd(counters + it.blk, 1u);
                if (P.npeers && old == (uint32_t)P.cfg.nchains_local - 1u) {   // this rank's block is complete: tell the peers
                  __threadfence_system();
                  for (int pz = 0; pz < P.npeers; ++pz) atomicMax_system(reinterpret_cast<unsigned long long *>(P.peer_flag[pz]), (unsigned long long)(P.ww_k0 + it.blk + 1));
                }
              } else peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
            }
          }
          if (g == 0 && valid) {
            *lrow_ptr = last_like + last_prior;
            // decision word (dreamzs_common.cuh pack_decision): snooker, CR index, gamma level, one DE pair, gamma == 1
            if (drow_ptr)
              *drow_ptr = (uint32_t)changed | ((mt >> 7) & 2u) | ((mt & 15u) << 2) | (((mt >> 4) & 15u) << 6) | (1u << 10) |
                          (((mt >> 9) & 1u) << 18) | ((uint32_t)accepted << 19);
          }
          trow_ptr += ld; lrow_ptr += 1; if (drow_ptr) drow_ptr += 1;
          WF_STAMP(0);   // c: iteration done (warp 0)
        }
        if ((!resident || it.last_window) && valid) {   // chain state -> global memory
          if (own) {
            double2 *xr = reinterpret_cast<double2 *>(P.st.X + (size_t)c_local * ld + i0);
            xr[0] = make_double2(x0[0], x0[1]); xr[1] = make_double2(x0[2], x0[3]);
            double2 *ur = reinterpret_cast<double2 *>(P.st.gauss_U + (size_t)c_local * ld + i0);
            ur[0] = make_double2(u0[0], u0[1]); ur[1] = make_double2(u0[2], u0[3]);
          }
          if (g == 0) { P.st.last_prior[c_local] = last_prior; P.st.last_like[c_local] = last_like; }
        }
      }
      // this warp is done with the item's slots (warps without chains in this group count as done as well)
      __threadfence_block();
      __syncwarp();
      if (lane == 0) st_vol(cfin + warp, it.seq + 1);
    }
  }
#undef WF_STAMP
}

struct WflowPlan { int tc, nb, lpc; size_t smem; WflowLayout layout; };

// chains per CTA for the dataflow kernel: whole windows only (nb = longest window), tc chains so that the C warps leave
// room for the M warps and at least two V warps
inline WflowPlan wflow_plan(const dreamzs_config &cfg, int sms, int wmax, int tc_force = 0) {
  WflowPlan pl{};
  const int ld = cfg.ld, nch = ld / 4;
  if (ld > 128 || (ld & 3) || ld < 8 || cfg.nchains_local < 1 || wmax < 1 || wmax > 32) return pl;
  pl.lpc = nch <= 8 ? 8 : nch <= 16 ? 16 : 32;
  const int cpw = 32 / pl.lpc;
  const size_t cap = 227 * 1024;
  int tc = (cfg.nchains_local + sms - 1) / sms;
  if (tc_force > 0) tc = tc_force;
  const int maxcw = (WW_WARPS - WF_NMW) / 2 + 1;     // at most about half of the remaining warps run chains
  if (tc > maxcw * cpw) tc = maxcw * cpw;
  for (; tc >= 1; --tc)
    if ((size_t)wflow_layout(cfg.ndim, ld, tc, wmax, cfg.ngamma).bytes <= cap && (tc * wmax + tc) * nch < (1 << 20) / 32) break;
  if (tc < 1) return pl;
  pl.tc = tc; pl.nb = wmax;
  pl.layout = wflow_layout(cfg.ndim, ld, tc, wmax, cfg.ngamma);
  pl.smem = (size_t)pl.layout.bytes;
  return pl;
}

template <int LPC>
int launch_wflow_t(StepParams &P, const WflowPlan &pl, int sms, cudaStream_t stream) {
  auto kern = dreamzs_wflow_kernel<LPC>;
  static size_t smem_set[64] = {0};
  if (ensure_dynamic_smem(kern, pl.smem, smem_set) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  const int ngroups = (P.cfg.nchains_local + pl.tc - 1) / pl.tc;
  int grid = ngroups;
  if (P.ww_sync) {
    // several windows: CTAs wait for rows other CTAs append, so all of them must be resident: at most one wave, a CTA
    // walks the chain groups grid-stride.  (A plain launch: the GPU may first finish other work before all CTAs are in.)
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WW_THREADS, pl.smem) != cudaSuccess || per_sm < 1) {
      (void)cudaGetLastError();
      return DREAMZS_E_LAUNCH;
    }
    const int cap = sms * per_sm;
    grid = ngroups < cap ? ngroups : cap;
  }
  kern<<<grid, WW_THREADS, pl.smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
