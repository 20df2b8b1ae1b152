// Whitened window kernel for the dense-Gaussian target (BASELINE configs C2 / C5), sm_100a.
//
// One launch = one window of <= history_thin iterations (the archive is constant inside it) for TC chains per CTA,
// 1024 threads.  A (chain, iteration) pair is a COLUMN.  Everything of Dream.astep that does not depend on the chain
// state -- all draws, the archive rows, the jump dx = (e*gamma)*(z_r1 - z_r2) + zeta with its crossover mask -- is a
// function of the Philox counters and of the archive, so it is produced for ALL columns of the window at once, flattened
// over the CTA's threads; only the Markov chain itself is serial.
//
// The quadratic form is carried in WHITENED coordinates: invC = L L^T (Cholesky factor computed on the host in extended
// precision), u = L^T x, Q(x) = |u|^2 and Q(x + dx) = |u + L^T dx|^2.  L is triangular, so the product costs d^2/2
// FMAs per column instead of d^2, and the chain needs no second dot product.  The products of all columns,
// DU = DX^T L, run on the fp64 tensor-core path (mma.sync m8n8k4 f64, SASS DMMA: same fp64 pipe as DFMA on B200 but
// 8x fewer issue slots), triangular tiles only.
//
// Phases (CTA-wide barriers in between; `task` loops are flattened over all 1024 threads):
//   S   scalar draws, one Philox block per (kind, column): snooker / CR / gamma level / gamma unity multinomials
//       (Dream.py:542-599, 615), the two np.random.uniform() calls (:618, :993), the random.sample calls (:646-668);
//       one thread per column assembles the decisions and stages the archive rows by TMA (cp.async.bulk, mbarrier
//       complete_tx) straight into the column's slots
//   V1  per (column, 4-dimension chunk): the crossover uniforms U (Dream.py:700) -> keep mask, d' (shared-memory atomic)
//   V2  per (column, chunk): zeta (float32 Box-Muller of the RNG contract) and e, gamma from d', then
//       J = (e*gamma)*(z_r1 - z_r2), dx = J + zeta in place of the staged rows (Dream.py:688-726)
//   M   DU = DX^T L per 8-column tile on DMMA, results held in registers until every column has been read, then
//       written in place of dx; a refresh window adds the chains' x as columns (u = L^T x)
//   C   the chains (LPC lanes per chain): per iteration prop = (x + J) + zeta, Q' = |u + du|^2 by ONE reduction, the
//       Metropolis test against the precomputed log u, state / trace / archive append.  The snooker move is linear in
//       the state: dx = c (x - z), du = c (u - L^T z), so its column of M is L^T z.
// RNG consumption, decisions and element-wise arithmetic are those of dreamzs_step_kernel / the oracle; only the
// summation order of the quadratic form differs.
#pragma once
#include "dreamzs_gauss_kernel.cuh"

namespace dreamzs {

#ifndef DZ_WW_THREADS
#define DZ_WW_THREADS 512
#endif
constexpr int WW_THREADS = DZ_WW_THREADS;   // A/B builds override (registers per thread = 65536 / threads)
constexpr int WW_WARPS = WW_THREADS / 32;
constexpr int WW_MAXI = WW_THREADS >= 1024 ? 8 : WW_THREADS >= 640 ? 10 : 16;   // i-tiles whose results one M warp holds in registers
constexpr int WW_MAXSPLIT = 4;    // i-ranges an 8-column tile is split into

__host__ __device__ inline int wwin_ntiles(int ld) {
  const int nK = ld / 4, nI = (ld + 7) / 8;
  int n = 0;
  for (int I = 0; I < nI; ++I) n += max(0, nK - 2 * I);
  return n;
}
// first tile of i-tile I in the packed factor: tiles (I, k = 2I .. nK-1), 32 doubles each in lane order
__host__ __device__ inline int wwin_tile0(int nK, int I) { return I * nK - I * (I - 1); }

// shared-memory carve-up (byte offsets), computed once on the host and passed with the launch parameters
inline WwinLayout wwin_layout(int d, int ld, int TC, int NB, int ngamma) {
  WwinLayout L;
  L.nch = ld / 4; L.nK = ld / 4; L.nI = (ld + 7) / 8;
  L.ntilesL = wwin_ntiles(ld);
  L.ncolmax = TC * NB;
  auto up = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t o = 0;
  L.oL = (int32_t)o;     o += up((size_t)L.ntilesL * 32 * 8);
  L.oW = (int32_t)o;     o += up((size_t)(L.ncolmax + TC) * ld * 8);     // dx / z columns -> du; + TC refresh columns
  L.oJ = (int32_t)o;     o += up((size_t)L.ncolmax * ld * 8);            // J = (e*gamma)*diff | snooker: z
  L.oN = (int32_t)o;     o += up((size_t)L.ncolmax * ld * 4);            // float32 normals of zeta (0 where the dimension is reset)
  L.oXs = (int32_t)o;    o += up((size_t)TC * ld * 8);
  L.oUs = (int32_t)o;    o += up((size_t)TC * ld * 8);
  L.oGam = (int32_t)o;   o += up((size_t)ngamma * d * 8);
  L.oScr = (int32_t)o;   o += up((size_t)9 * L.ncolmax * 8);             // raw scalar draws [kind][column]
  L.oLogu = (int32_t)o;  o += up((size_t)2 * L.ncolmax * 8);             // from here: two sets (batch parity)
  L.oGsn = (int32_t)o;   o += up((size_t)2 * L.ncolmax * 8);
  L.oRows = (int32_t)o;  o += up((size_t)2 * L.ncolmax * 24);            // row indices: DE r1, r2 | snooker z, z1, z2
  L.oMeta = (int32_t)o;  o += up((size_t)2 * L.ncolmax * 4);
  L.oDpr = (int32_t)o;   o += up((size_t)2 * L.ncolmax * 4);
  L.oMask = (int32_t)o;  o += up((size_t)2 * L.ncolmax * L.nch);
  L.oMbar = (int32_t)o;  o += up((size_t)(L.ncolmax + 1) * 8);
  L.oUses = (int32_t)o;  o += up((size_t)L.ncolmax);                     // TMA uses of a column slot (mbarrier phase)
  L.oProbs = (int32_t)o; o += 48 * 8;                // [0,16) CR, [16,24) gamma level, 24 snooker, 26 unity, 32 abort flag
  L.oCst = (int32_t)o;   o += up((size_t)TC * 4 * 8);
  // pool of z1 - z2 rows for snooker columns (filled in V2, read by the chains): whatever shared memory is left, at most
  // one slot per column; columns that find the pool full read z1, z2 from the archive inside the chain loop
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

// x / div for x * div < 2^20 with m = ceil(2^20 / div) (exact in that range): the flattened task loops split a task
// number into (column, chunk) and a column into (chain, iteration) without integer division
__device__ __forceinline__ int fdiv20(int x, uint32_t m) { return (int)(((uint32_t)x * m) >> 20); }
__device__ __forceinline__ uint32_t fdiv20_magic(int div) { return ((1u << 20) + (uint32_t)div - 1u) / (uint32_t)div; }

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Philox4x32-10 inlined (the flattened V passes are straight-line code; two independent blocks interleave)
__device__ __forceinline__ uint4 philox_inl(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Sum over the LPC lanes of a chain on the fp64 MMA path, two DMMAs instead of log2(LPC) shuffle rounds (the chain phase is
// bound by dependent latency and the fp64 pipe is idle there).  With A = ones, D = A B sums the B fragment's columns:
// lane l gets the sums of lanes {8m..8m+3} and {8m+4..8m+7}, m = l % 4; their sum T[m] goes through a second product
// whose A fragment selects the T's of the lane's own chain.  Every lane of a chain ends with the same bits.
template <int LPC> __device__ __forceinline__ double lsum(double v);
template <int LPC>
__device__ __forceinline__ double lsum_mma(double v, double a2) {
#ifdef DZ_WW_SHFL_SUM   // A/B: butterfly of shuffles instead
  return lsum<LPC>(v);
#endif
  double d0 = 0.0, d1 = 0.0;
  dmma884(d0, d1, 1.0, v);
  const double t = d0 + d1;
  double e0 = 0.0, e1 = 0.0;
  dmma884(e0, e1, a2, t);
  return e0;
}
// the second product's A fragment: row i = lane / 4 (a lane of chain i * 4 / LPC), column k = lane % 4 (T[k] = lanes 8k..8k+7)
template <int LPC>
__device__ __forceinline__ double lsum_mma_sel(int lane) {
  return ((lane >> 2) * 4) / LPC == ((lane & 3) * 8) / LPC ? 1.0 : 0.0;
}

template <int LPC>
__device__ __forceinline__ double lsum(double v) {
#pragma unroll
  for (int o = LPC / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// the row is on its way from HBM to L2 (sm_90+ bulk prefetch): issued a batch ahead of the TMA copy that stages it
__device__ __forceinline__ void l2_prefetch_row(const double *row, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"(bytes) : "memory");
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// The peers' flags live in THIS GPU's memory and so do the rows they announce (the writer fences at system scope between
// the two): for the reader a gpu-scope acquire is what orders its later reads of the rows -- through this GPU's L2 --
// after the flag.  (ld.acquire.sys costs a system-scope fence per poll: ~14 us measured on B200, which put the pre-draw
// warps behind the chains at 2 GPUs.)
__device__ __forceinline__ uint64_t ld_acquire_gpu_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Where can archive row r be read?  Rows below the launch's archive size: in this replica.  A row appended DURING this
// launch belongs to append block j = (r - archive_rows) / nchains_global and to the chain (r - archive_rows) % nchains_global:
//   - a local chain: the row is there once the chains of its group (the TC chains that share a CTA) have made append j
//     (gdone[g]: every chain counts itself after its row is visible on this GPU), or -- without per-group words -- once
//     every local chain has (counters[j]);
//   - a chain of peer q: in this replica once q has confirmed the block (flags, known[1 + q]); before that the row is read
//     from q's OWN archive over NVLink as soon as q's group has completed the append (peer_pub, polled in q's memory).
// Gives up after DREAMZS_PEER_TIMEOUT_NS or when another CTA has aborted.
struct RowWait {      // what row_source needs of the launch parameters (by value: the parameter block stays in constant memory)
  int64_t archive_rows;
  int32_t nchains_global, chain_begin, nchains_local;
  uint64_t k0;
  const uint64_t *my_flags;
  int32_t *peer_error;
  const uint32_t *counters;
  volatile int32_t *status;
  long long *dbg;       // profiling aid: [66] cycles spent waiting for rows, [67] waits
  const uint32_t *gdone;
  int32_t tc, rank;
  const uint64_t *peer_pub[DREAMZS_MAX_PEERS];
};
constexpr int ROW_SRC_SHIFT = 48;      // a row index carries its source above bit 48: 0 this replica, 1 + pz peer pz's archive
__device__ __noinline__ bool row_wait(const RowWait w, int j, bool local, int q, int grp, const uint64_t *pub) {
  const uint64_t k = w.k0 + (uint64_t)j + 1u;
  const uint32_t need = (uint32_t)(j + 1) * (uint32_t)min(w.tc, w.nchains_local - grp * w.tc);   // (local groups)
  const uint64_t t0 = globaltimer_ns();
  for (;;) {
    if (pub) { if (ld_relaxed_sys_u64(pub + grp) >= k) return true; }
    else if (local && w.gdone) { if (ld_acquire_gpu_u32(w.gdone + grp) >= need) return true; }
    else if (local ? ld_acquire_gpu_u32(w.counters + j) >= (uint32_t)w.nchains_local : ld_acquire_gpu_u64(w.my_flags + q) >= k) return true;
    if (*w.status != 0) return false;
    if (globaltimer_ns() - t0 > DREAMZS_PEER_TIMEOUT_NS) {
      atomicExch(const_cast<int32_t *>(w.status), 1);
      if (w.peer_error) atomicExch(w.peer_error, 1);
      return false;
    }
    __nanosleep(32);
  }
}
// `known` (shared memory, zeroed per launch): known[0] = leading append blocks of this launch known complete for the local
// chains, known[1 + q] = the same for peer q's rows in this replica (brought up to date once per batch by known_refresh).
// Returns -1 (timed out / aborted), 0 (this replica) or 1 + pz (peer pz's archive).
__device__ __forceinline__ int row_source(const RowWait &w, int *known, int64_t r) {
  if (r < w.archive_rows || !w.counters) return 0;
  const int64_t off = r - w.archive_rows;
  const int j = (int)(off / w.nchains_global);
  const int owner = (int)(off - (int64_t)j * w.nchains_global);
  const bool local = owner >= w.chain_begin && owner < w.chain_begin + w.nchains_local;
  const int q = local ? 0 : owner / w.nchains_local;
  int *kn = known + (local ? 0 : 1 + q);
  if (j < *reinterpret_cast<volatile int *>(kn)) return 0;
  const int pz = q < w.rank ? q : q - 1;
  const uint64_t *pub = local ? nullptr : w.peer_pub[pz];
  const int grp = w.tc > 0 ? (owner - (local ? w.chain_begin : q * w.nchains_local)) / w.tc : 0;
  const long long t0 = w.dbg ? clock64() : 0;
  if (!row_wait(w, j, local, q, grp, pub)) return -1;
  if (w.dbg) { atomicAdd(reinterpret_cast<unsigned long long *>(w.dbg) + 66, (unsigned long long)(clock64() - t0)); atomicAdd(reinterpret_cast<unsigned long long *>(w.dbg) + 67, 1ull); }
  if (pub) { __threadfence(); return 1 + pz; }
  if (!(local && w.gdone)) atomicMax(kn, j + 1);     // whole-block wait: everything up to j is there
  return 0;
}
// true when row r lies in a block appended during this launch that is not known complete yet (shared-memory test only)
__device__ __forceinline__ bool row_maybe_late(const RowWait &w, const int *known, int64_t r) {
  if (r < w.archive_rows || !w.counters) return false;
  const int64_t off = r - w.archive_rows;
  const int j = (int)(off / w.nchains_global);
  const int owner = (int)(off - (int64_t)j * w.nchains_global);
  const bool local = owner >= w.chain_begin && owner < w.chain_begin + w.nchains_local;
  return j >= *reinterpret_cast<const volatile int *>(known + (local ? 0 : 1 + owner / w.nchains_local));
}
// one thread, once per batch: how many leading blocks of this launch are complete by now (blocks [0, nblk) exist)
__device__ __forceinline__ void known_refresh(const RowWait &w, int *known, int nblk, int world, int rank) {
  if (!w.counters) return;
  int k = known[0];
  while (k < nblk && ld_acquire_gpu_u32(w.counters + k) >= (uint32_t)w.nchains_local) ++k;
  if (k > known[0]) atomicMax(known, k);
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    const uint64_t f = ld_acquire_gpu_u64(w.my_flags + q);
    const int kq = f > w.k0 ? (int)min((uint64_t)nblk, f - w.k0) : 0;
    if (kq > known[1 + q]) atomicMax(known + 1 + q, kq);
  }
}
#ifndef DZ_WW_MINBLOCKS
#define DZ_WW_MINBLOCKS 1
#endif

// one batch of one window of one group of chains: the unit the kernel's main loop works on
struct WwItem {
  int64_t wt0, nxt, M, trace_row0;
  int wphase, blk, grp, done, seq;
  int wn, nb, ncol, cta_chain0, nch_cta;
  bool aligned, first_window, valid, w_append, w_refresh, last_window, do_refresh, first_batch, last_batch;
};
__device__ __forceinline__ void ww_derive(WwItem &w, const StepParams &P, int TC, int NB) {
  const int64_t t_end = P.iter_begin + P.niter;
  w.valid = w.wt0 < t_end;
  if (!w.valid) return;
  w.wn = (int)((t_end < w.nxt + 1 ? t_end : w.nxt + 1) - w.wt0);         // iterations of this window
  w.w_append = w.wt0 + w.wn - 1 == w.nxt;
  w.w_refresh = w.wt0 == 0 || (w.aligned && w.wphase == 0);
  w.last_window = w.wt0 + w.wn >= t_end;
  w.M = P.archive_rows + (int64_t)w.blk * P.cfg.nchains_global;          // archive rows this window samples
  w.trace_row0 = P.tr.trace_offset + (w.wt0 - P.iter_begin);
  w.cta_chain0 = w.grp * TC;
  w.nch_cta = min(TC, P.cfg.nchains_local - w.cta_chain0);               // chains of this group
  w.nb = min(NB, w.wn - w.done);
  w.ncol = w.nch_cta * w.nb;                                             // columns of the batch: col = chain * nb + iteration
  w.first_batch = w.done == 0;
  w.last_batch = w.done + w.nb >= w.wn;
  w.do_refresh = w.w_refresh && w.first_batch;
}
__device__ __forceinline__ void ww_first(WwItem &w, const StepParams &P, int TC, int NB) {
  const int64_t thin = P.cfg.history_thin;
  w.wt0 = P.iter_begin;
  w.nxt = ((P.iter_begin + thin - 1) / thin) * thin;          // first appending iteration >= the window's first
  // a window that starts right after append #a (a = (wt0 - 1) / thin) re-derives u from x when a % REFRESH == 0
  w.wphase = (int)((P.iter_begin > 0 ? (P.iter_begin - 1) / thin : 0) % DREAMZS_GAUSS_REFRESH_WINDOWS);
  w.aligned = P.iter_begin == 0 || (P.iter_begin - 1) % thin == 0;
  w.blk = 0; w.grp = blockIdx.x; w.done = 0; w.seq = 0; w.first_window = true;
  ww_derive(w, P, TC, NB);
}
__device__ __forceinline__ void ww_next(WwItem &w, const StepParams &P, int TC, int NB, int ngroups, int nwork) {
  w.seq += 1;
  w.done += w.nb;
  if (w.done >= w.wn) {            // next group of the window; after the last one the next window
    w.done = 0;
    w.grp += nwork;
    if (w.grp >= ngroups) {
      w.grp = blockIdx.x;
      w.first_window = false;
      if (w.w_append) {
        w.blk += 1;
        if (w.nxt > 0) w.wphase = w.wphase + 1 == DREAMZS_GAUSS_REFRESH_WINDOWS ? 0 : w.wphase + 1;   // the next window follows append #(a + 1)
        w.nxt += P.cfg.history_thin;
        w.aligned = true;
      } else w.aligned = false;
      w.wt0 += w.wn;
    }
  }
  ww_derive(w, P, TC, NB);
}

template <int LPC>
__global__ void __launch_bounds__(WW_THREADS, DZ_WW_MINBLOCKS) dreamzs_wwin_kernel(const StepParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = P.cfg.ndim, ld = P.cfg.ld, TC = P.ww_tc, NB = P.ww_nb;
  const WwinLayout &L = P.ww_L;
  const int nch = L.nch, nK = L.nK, NCM = L.ncolmax;
  double *Lf = reinterpret_cast<double *>(smem_raw + L.oL), *Wc = reinterpret_cast<double *>(smem_raw + L.oW);
  double *Jc = reinterpret_cast<double *>(smem_raw + L.oJ);
  float *Nz = reinterpret_cast<float *>(smem_raw + L.oN);
  double *Xs = reinterpret_cast<double *>(smem_raw + L.oXs), *Us = reinterpret_cast<double *>(smem_raw + L.oUs);
  double *gam = reinterpret_cast<double *>(smem_raw + L.oGam);
  uint2 *scr = reinterpret_cast<uint2 *>(smem_raw + L.oScr);
  // two sets of the per-column scalars (batch parity): the draws of the NEXT batch are made while the chains run this one
  double *logu2 = reinterpret_cast<double *>(smem_raw + L.oLogu), *gsn2 = reinterpret_cast<double *>(smem_raw + L.oGsn);
  int64_t *rows2 = reinterpret_cast<int64_t *>(smem_raw + L.oRows);
  uint32_t *meta2 = reinterpret_cast<uint32_t *>(smem_raw + L.oMeta);
  int *dpr2 = reinterpret_cast<int *>(smem_raw + L.oDpr);
  unsigned char *mask2 = smem_raw + L.oMask;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + L.oMbar);   // [0, ncolmax) columns, [ncolmax] the factor
  unsigned char *uses = smem_raw + L.oUses;                            // TMA uses of a column slot so far (mbarrier phase)
  double *probs = reinterpret_cast<double *>(smem_raw + L.oProbs), *cst = reinterpret_cast<double *>(smem_raw + L.oCst);
  double *pool = reinterpret_cast<double *>(smem_raw + L.oPool);
  int *pool_n = reinterpret_cast<int *>(probs + 34);          // snooker columns of the batch that hold a pool slot
  int *known = reinterpret_cast<int *>(probs + 35);           // [1 + DREAMZS_MAX_PEERS] append blocks known complete (row_ready)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double logF = P.st.target_table[0];
  int dbg_n = 0;
#define WW_STAMP() do { if (P.dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 60) P.dbg[dbg_n] = clock64(); ++dbg_n; } while (0)
  WW_STAMP();   // 0: kernel entry
  if (P.ww_confirm && blockIdx.x == gridDim.x - 1) {
    // ---- the confirmer: several GPUs, several windows per launch.  The chains push their appended rows to the peers'
    // replicas and fence at GPU scope only; for every append of the launch this one thread waits until all local chains
    // have counted themselves, fences ONCE at system scope (~14 us on B200 -- on nobody's path here) and tells every peer
    // "block #k of this rank is in your replica".  Until then the peers fetch those rows from this rank's archive.
    if (threadIdx.x == 0) {
      const uint32_t *cnt = P.ww_sync + 16;
      const volatile int32_t *st = reinterpret_cast<const volatile int32_t *>(P.ww_sync);
      const int napp = P.ww_wcap - 2;
      for (int j = 0; j < napp; ++j) {
        const uint64_t t0 = globaltimer_ns();
        while (ld_acquire_gpu_u32(cnt + j) < (uint32_t)P.cfg.nchains_local) {
          if (*st != 0 || globaltimer_ns() - t0 > 4 * DREAMZS_PEER_TIMEOUT_NS) return;
          __nanosleep(200);
        }
        __threadfence_system();
        for (int pz = 0; pz < P.npeers; ++pz)
          atomicMax_system(reinterpret_cast<unsigned long long *>(P.peer_flag[pz]), (unsigned long long)(P.ww_k0 + (uint64_t)j + 1u));
      }
    }
    return;
  }

  const uint32_t row_bytes = (uint32_t)ld * 8u;
  const uint32_t s0 = P.cfg.snooker != 0 ? 1u : 0u;   // multinomial call number of the CR draw
  const uint32_t k0 = (uint32_t)P.cfg.seed, k1 = (uint32_t)(P.cfg.seed >> 32);
  const int ngroups = (P.cfg.nchains_local + TC - 1) / TC;          // groups of TC chains; a CTA takes groups blockIdx.x, + gridDim.x, ...
  const int nwork = (int)gridDim.x - (P.ww_confirm ? 1 : 0);       // CTAs that walk chain groups (the last one may be the confirmer)
  const bool resident = ngroups <= nwork;                           // one group per CTA: chain states stay in shared memory
  volatile int32_t *status = reinterpret_cast<volatile int32_t *>(P.ww_sync);     // word 0: != 0 aborts the launch
  uint32_t *counters = P.ww_sync ? P.ww_sync + 16 : nullptr;                     // chains that have made append #j of this launch
  constexpr int CPW = 32 / LPC;                                     // chains per warp in the chain phase
  const int NCW = (TC + CPW - 1) / CPW;                             // warps [0, NCW) run the chains
  const int n_pre = (WW_WARPS - NCW) * 32;                          // threads that draw the next batch meanwhile
  const bool overlap = n_pre >= 64;                                 // (too few left: the draws follow the chains instead)

  // ---- prologue: one TMA bulk copy brings the packed factor; tables -> shared memory
  if (tid <= NCM) mbar_init(mbar + tid, 1);
  if (tid < 48) {
    double v = 0.0;
    if (tid < 16) v = tid < P.cfg.nCR ? P.st.cr_probs[tid] : 0.0;
    else if (tid < 24) v = tid - 16 < P.cfg.ngamma ? P.st.gamma_probs[tid - 16] : 0.0;
    else if (tid == 24) v = P.cfg.snooker;
    else if (tid == 26) v = P.cfg.p_gamma_unity;
    probs[tid] = v;
  }
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
    probs[32] = bad ? 1.0 : 0.0;
  }
  __syncthreads();
  if (probs[32] != 0.0) return;
  if (tid == 0) {
    fence_proxy_async();
    const uint32_t bytes = (uint32_t)L.ntilesL * 256u;
    mbar_expect_tx(mbar + NCM, bytes);
    tma_load_row(Lf, P.st.gauss_L, bytes, mbar + NCM);
  }
  WW_STAMP();   // 1: prologue done
  bool factor_ready = false;
  long long tprev = (P.dbg && tid == 0) ? clock64() : 0;
  RowWait rw = {P.archive_rows, P.cfg.nchains_global, P.cfg.chain_begin, P.cfg.nchains_local, P.ww_k0, P.my_flags,
                P.peer_error, counters, status, P.dbg, P.ww_gdone, TC, P.my_rank, {}};
  for (int pz = 0; pz < DREAMZS_MAX_PEERS; ++pz) rw.peer_pub[pz] = (P.my_pub && pz < P.npeers) ? P.peer_pub[pz] : nullptr;
  // the archive a (tagged) row index refers to: this replica, or -- for a peer's row not yet confirmed here -- the owner's
  auto zrow = [&](int64_t r) -> const double * {
    const int src = (int)(r >> ROW_SRC_SHIFT);
    return (src ? P.peer_Z[src - 1] : P.st.Z) + (size_t)(r & ((1ll << ROW_SRC_SHIFT) - 1)) * ld;
  };

  // ================================================================ pre(batch): everything of a batch that needs neither
  // the column slots nor the chain state -- scalar draws, decisions, archive row indices (waiting for rows appended inside
  // this launch), crossover masks and d' -- by the threads [ptid of pn], synchronised on named barrier `bar`
  auto pre = [&](const WwItem &w, int ptid, int pn, int bar) {
    const int set = w.seq & 1;
    double *logu = logu2 + set * NCM, *gsn = gsn2 + set * NCM;
    int64_t *rows = rows2 + (size_t)set * NCM * 3;
    uint32_t *meta = meta2 + set * NCM;
    int *dpr = dpr2 + set * NCM;
    unsigned char *maskb = mask2 + (size_t)set * NCM * nch;
    const int ncol = w.ncol, nb = w.nb;
    const uint32_t m_nb = fdiv20_magic(nb), m_nch = L.m_nch;
    // ---- S: scalar draws (thread per (kind, column)), one Philox block each:
    //      0-3: snooker, CR, gamma level, gamma unity (Dream.py:542-599, 615); 4-5: the first two np.random.uniform()
    //      (snooker gamma :618 / Metropolis :993); 6-8: random.sample calls 0-2 (sample_from_history, :646-668)
    for (int task = ptid; task < 9 * ncol; task += pn) {
      int kind = 0, col = task;
      while (col >= ncol) { col -= ncol; ++kind; }
      const int ch = fdiv20(col, m_nb), itb = col - ch * nb;
      const uint32_t iter = (uint32_t)(w.wt0 + w.done + itb);
      const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + w.cta_chain0 + ch);
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
    if (ptid == 0) *pool_n = 0;
    if (ptid == pn - 1) known_refresh(rw, known, w.blk, P.my_flags ? P.world : 0, P.my_rank);
    named_sync(bar, pn);
    // ---- one thread per column: decisions, archive rows
    for (int col = ptid; col < ncol; col += pn) {
      const uint2 q0 = scr[col], q1 = scr[NCM + col], q2 = scr[2 * NCM + col], q3 = scr[3 * NCM + col];
      const uint2 u4 = scr[4 * NCM + col], u5 = scr[5 * NCM + col];
      const uint2 r6 = scr[6 * NCM + col], r7 = scr[7 * NCM + col], r8 = scr[8 * NCM + col];
      const bool snk = (s0 != 0u) && q0.x == 0u;
      const uint32_t use = uses[col];
      uses[col] = (unsigned char)(use + 1u);
      // meta word: bits 0-3 CR index, 4-7 gamma level, 8 snooker, 9 gamma == 1 (set in V2), 10 "not unity",
      // 11 parity of the mbarrier phase this use of the column slot completes, 12 a sampled row was appended inside this
      // launch and is not known to be there yet (the thread that stages the column waits for it; V2 takes the column last)
      uint32_t mt = q1.x | (q2.x << 4) | (snk ? 256u : 0u) | (q3.x != 0u ? 1024u : 0u) | ((use & 1u) << 11);
      // Metropolis uniform: the 2nd np.random.uniform() after a snooker gamma, else the 1st (its log is in place)
      if (snk) logu[col] = __hiloint2double((int)u5.y, (int)u5.x);
      if (!snk) {
        const int64_t ra = (int64_t)(((uint64_t)r6.x * (uint64_t)w.M) >> 32);
        int64_t rb = (int64_t)(((uint64_t)r6.y * (uint64_t)(w.M - 1)) >> 32);
        if (rb >= ra) rb += 1;
        rows[3 * col] = ra; rows[3 * col + 1] = rb; rows[3 * col + 2] = -1;
        if (row_maybe_late(rw, known, ra) || row_maybe_late(rw, known, rb)) mt |= 4096u;
        else { l2_prefetch_row(P.st.Z + (size_t)ra * ld, row_bytes); l2_prefetch_row(P.st.Z + (size_t)rb * ld, row_bytes); }
        dpr[col] = 0;
        gsn[col] = ((double)(q1.x + 1u) / (double)P.cfg.nCR) * 4294967296.0;   // DE column: CR 2^32 for the crossover test
      } else {
        const double g = 1.2 + (2.2 - 1.2) * u53_of(u4.x, u4.y);             // snooker gamma, Dream.py:618
        gsn[col] = g;
        if (g == 1.0) mt |= 512u;
        const int64_t rz = (int64_t)(((uint64_t)r6.x * (uint64_t)w.M) >> 32);
        const int64_t r1 = (int64_t)(((uint64_t)r7.x * (uint64_t)w.M) >> 32), r2 = (int64_t)(((uint64_t)r8.x * (uint64_t)w.M) >> 32);
        rows[3 * col] = rz; rows[3 * col + 1] = r1; rows[3 * col + 2] = r2;
        if (row_maybe_late(rw, known, rz) || row_maybe_late(rw, known, r1) || row_maybe_late(rw, known, r2)) mt |= 4096u;
        else {
          l2_prefetch_row(P.st.Z + (size_t)rz * ld, row_bytes); l2_prefetch_row(P.st.Z + (size_t)r1 * ld, row_bytes);
          l2_prefetch_row(P.st.Z + (size_t)r2 * ld, row_bytes);
        }
        const int slot = atomicAdd(pool_n, 1);                               // z1 - z2 goes to the pool when a slot is left
        dpr[col] = slot < L.npool ? slot : -1;
      }
      meta[col] = mt;
    }
    named_sync(bar, pn);
    // ---- V1: crossover uniforms -> keep mask, d'
    const int ntask = ncol * nch;
    for (int task = ptid; task < ntask; task += pn) {
      const int col = fdiv20(task, m_nch), q = task - col * nch;
      const uint32_t mt = meta[col];
      if (mt & 256u) continue;
      const int ch = fdiv20(col, m_nb), itb = col - ch * nb;
      const uint32_t iter = (uint32_t)(w.wt0 + w.done + itb);
      const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + w.cta_chain0 + ch);
      const uint4 wu = philox_inl((uint32_t)q, (1u << 3) | ST_UNIFORM_VEC, iter, c_global, k0, k1);
      // U = w 2^-32 exactly, so U < CR <=> w < ceil(CR 2^32) and U > CR <=> w > floor(CR 2^32)
      const double CRs = gsn[col];
      const uint64_t t_lt = (uint64_t)ceil(CRs), t_gt = (uint64_t)floor(CRs);
      const bool lt_all = (t_lt >> 32) != 0, gt_none = (t_gt >> 32) != 0;   // CR = 1: every U < CR, none > CR
      const uint32_t tl = (uint32_t)t_lt, tg = (uint32_t)t_gt;
      const uint32_t wv[4] = {wu.x, wu.y, wu.z, wu.w};
      unsigned reset = 0;
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (4 * q + j < d) {
          cnt += (lt_all || wv[j] < tl);
          if (!gt_none && wv[j] > tg) reset |= 1u << j;
        } else reset |= 1u << j;
      }
      maskb[task] = (unsigned char)reset;
      if (cnt) atomicAdd(dpr + col, cnt);
    }
  };

  // ================================================================ batches of the launch (pydream/core.py:103-122)
  // A window ends at an appending iteration (t % history_thin == 0); the archive it samples is the one of the launch's
  // start plus the appends of the windows before it.  CTAs do not synchronise per window: a column whose sampled row was
  // appended during THIS launch waits for the chains that write that block (counters / peer flags), nothing else waits.
  WwItem it;
  ww_first(it, P, TC, NB);
  if (it.valid) pre(it, tid, WW_THREADS, 1);
  __syncthreads();
  while (it.valid) {
    if (probs[32] != 0.0) return;   // a wait for appended rows timed out (or another CTA aborted): give up, the host raises
    const int set = it.seq & 1;
    double *logu = logu2 + set * NCM, *gsn = gsn2 + set * NCM;
    int64_t *rows = rows2 + (size_t)set * NCM * 3;
    uint32_t *meta = meta2 + set * NCM;
    int *dpr = dpr2 + set * NCM;
    unsigned char *maskb = mask2 + (size_t)set * NCM * nch;
    const int64_t wt0 = it.wt0, M = it.M, trace_row0 = it.trace_row0;
    const int done = it.done, nb = it.nb, ncol = it.ncol, wn = it.wn, blk = it.blk;
    const int cta_chain0 = it.cta_chain0, nch_cta = it.nch_cta;
    const bool w_append = it.w_append, do_refresh = it.do_refresh;
    const uint32_t m_nb = fdiv20_magic(nb), m_nch = L.m_nch;
    const int ntask = ncol * nch;
    if (it.first_batch && (!resident || it.first_window)) {   // chain states -> shared memory
      for (int i = tid; i < nch_cta * nch; i += WW_THREADS) {
        const int cs = fdiv20(i, L.m_nch), q = i - cs * nch;
        const int c_local = cta_chain0 + cs;
        const double2 *xr = reinterpret_cast<const double2 *>(P.st.X + (size_t)c_local * ld + 4 * q);
        double2 *xd = reinterpret_cast<double2 *>(Xs + (size_t)cs * ld + 4 * q);
        xd[0] = xr[0]; xd[1] = xr[1];
        if (!it.w_refresh) {
          const double2 *ur = reinterpret_cast<const double2 *>(P.st.gauss_U + (size_t)c_local * ld + 4 * q);
          double2 *ud = reinterpret_cast<double2 *>(Us + (size_t)cs * ld + 4 * q);
          ud[0] = ur[0]; ud[1] = ur[1];
        }
      }
      if (tid < nch_cta) {
        cst[tid * 4 + 1] = P.st.last_prior[cta_chain0 + tid];
        cst[tid * 4 + 2] = P.st.last_like[cta_chain0 + tid];
      }
    }
    // ---- the archive rows of the batch's columns, staged by TMA straight into the column slots:
    //      DE: z_r1 -> J slot, z_r2 -> W slot;   snooker: z -> J slot and W slot (-> L^T z)
    // A column whose rows were appended inside this launch and may not have arrived (meta bit 12) is staged as soon as
    // the chains that write them -- other CTAs, other GPUs -- have: its thread waits, everybody else goes on (V2 takes such
    // columns last), so the latency of the append -> sample dependency is covered by the batch's other columns.
    for (int col = tid; col < ncol; col += WW_THREADS) {
      const uint32_t mt = meta[col];
      int64_t r3[3] = {rows[3 * col], rows[3 * col + 1], rows[3 * col + 2]};
      if (mt & 4096u) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (r3[i] >= 0) {
            const int src = row_source(rw, known, r3[i]);
            if (src < 0) probs[32] = 1.0;   // timed out: the host raises
            else if (src > 0) { r3[i] |= (int64_t)src << ROW_SRC_SHIFT; rows[3 * col + i] = r3[i]; }
          }
      }
      fence_proxy_async();   // the slots' earlier generic-proxy accesses and the acquired rows are ordered before the async copies
      mbar_expect_tx(mbar + col, 2u * row_bytes);
      const int64_t ra = r3[0], rb = (mt & 256u) ? ra : r3[1];
      tma_load_row(Jc + (size_t)col * ld, zrow(ra), row_bytes, mbar + col);
      tma_load_row(Wc + (size_t)col * ld, zrow(rb), row_bytes, mbar + col);
    }
    WW_STAMP();   // +0: rows requested
    const long long tp0 = (P.dbg && tid == 0) ? clock64() : 0;
    // ================================================================ V2: zeta, e, gamma -> J, dx in place of the rows
    for (int pass = 0; pass < 2; ++pass)
    for (int task = tid; task < ntask; task += WW_THREADS) {
      const int col = fdiv20(task, m_nch), q = task - col * nch;
      const uint32_t mt = meta[col];
      if (((mt >> 12) & 1u) != (uint32_t)pass) continue;      // second pass: the columns that may have had to wait
      if (mt & 256u) {   // snooker column: z1 - z2 (Dream.py:810) -> pool slot; z / L^T z arrive by TMA
        if (pass) mbar_wait(mbar + col, (mt >> 11) & 1u);     // (the staging thread has seen all three rows by then)
        const int slot = dpr[col];
        if (slot >= 0) {
          // (L2 loads: a row appended during this launch must not come from a stale L1 line it shares with its neighbour)
          const double *z1 = zrow(rows[3 * col + 1]) + 4 * q, *z2 = zrow(rows[3 * col + 2]) + 4 * q;
          const double2 p01 = __ldcg(reinterpret_cast<const double2 *>(z1)), p23 = __ldcg(reinterpret_cast<const double2 *>(z1) + 1);
          const double2 q01 = __ldcg(reinterpret_cast<const double2 *>(z2)), q23 = __ldcg(reinterpret_cast<const double2 *>(z2) + 1);
          double2 *bs = reinterpret_cast<double2 *>(pool + (size_t)slot * ld + 4 * q);
          bs[0] = make_double2(p01.x - q01.x, p01.y - q01.y);
          bs[1] = make_double2(p23.x - q23.x, p23.y - q23.y);
        }
        if (!pass) mbar_wait(mbar + col, (mt >> 11) & 1u);
        continue;
      }
      const int ch = fdiv20(col, m_nb), itb = col - ch * nb;
      const uint32_t iter = (uint32_t)(wt0 + done + itb);
      const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + cta_chain0 + ch);
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
      double *js = Jc + (size_t)col * ld + 4 * q, *ws = Wc + (size_t)col * ld + 4 * q;
      float *ns = Nz + (size_t)col * ld + 4 * q;
      mbar_wait(mbar + col, (mt >> 11) & 1u);
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
    if (do_refresh) {   // refresh columns: x of every chain, right after the batch's columns
      for (int i = tid; i < nch_cta * nch; i += WW_THREADS) {
        const int cs = fdiv20(i, m_nch), q = i - cs * nch;
        double2 *wd = reinterpret_cast<double2 *>(Wc + (size_t)(ncol + cs) * ld + 4 * q);
        const double2 *xs = reinterpret_cast<const double2 *>(Xs + (size_t)cs * ld + 4 * q);
        wd[0] = xs[0]; wd[1] = xs[1];
      }
    }
    if (!factor_ready) { mbar_wait(mbar + NCM, 0); factor_ready = true; }   // the factor has landed
    __syncthreads();
    WW_STAMP();   // +1: columns generated
    const long long tp1 = (P.dbg && tid == 0) ? clock64() : 0;
    // ================================================================ M: DU = DX^T L on DMMA
    // unit = (8-column tile, range of i-tiles) -> one warp (the host picks TC, NB and the split so that units <= 32 and a
    // range holds <= WW_MAXI i-tiles).  D fragment: lane l holds du[column l/4][8 I + 2 (l%4) + {0,1}].
    {
      const int ncols_total = ncol + (do_refresh ? nch_cta : 0);
      const int NT = (ncols_total + 7) / 8;
      const int nsplit = P.ww_nsplit;
      const int l4 = lane >> 2, lm = lane & 3;
      const bool mwarp = warp < NT * nsplit;
      const int nt = warp / nsplit, sp = warp - nt * nsplit;
      const int Ia = P.ww_isplit[mwarp ? sp : 0], Ib = P.ww_isplit[mwarp ? sp + 1 : 0];
      double acc[WW_MAXI][2];
#pragma unroll
      for (int t = 0; t < WW_MAXI; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
      if (mwarp) {
        const int crow = min(nt * 8 + l4, ncols_total - 1);     // padding rows alias the last column (results dropped)
        const double *arow = Wc + (size_t)crow * ld + lm;
#pragma unroll
        for (int t = 0; t < WW_MAXI; t += 2) {
          const int I0 = Ia + t, I1 = I0 + 1;
          if (I0 < Ib) {
            const bool two = I1 < Ib;
            const int ka = 2 * I0, kb = two ? min(2 * I1, nK) : nK;   // k in [ka, kb) feeds I0 only, [kb, nK) feeds both
            const double *b0 = Lf + (size_t)wwin_tile0(nK, I0) * 32 + lane;
            const double *b1 = Lf + (size_t)wwin_tile0(nK, two ? I1 : I0) * 32 + lane;
            const double *ap = arow + 4 * ka;
#pragma unroll 2
            for (int k = ka; k < kb; ++k, ap += 4, b0 += 32) dmma884(acc[t][0], acc[t][1], *ap, *b0);
            if (two) {
#pragma unroll 4
              for (int k = kb; k < nK; ++k, ap += 4, b0 += 32, b1 += 32) {
                const double a = *ap;
                dmma884(acc[t][0], acc[t][1], a, *b0);
                dmma884(acc[t + 1][0], acc[t + 1][1], a, *b1);
              }
            }
          }
        }
      }
      __syncthreads();   // every column has been read: the products may now overwrite dx in place
      WW_STAMP();   // +3: products computed
      if (mwarp) {
        const int c = nt * 8 + l4;
        if (c < ncols_total) {
          double *wrow = Wc + (size_t)c * ld + 2 * lm;
#pragma unroll
          for (int t = 0; t < WW_MAXI; ++t) {
            const int I = Ia + t;
            if (I < Ib && 8 * I + 2 * lm < ld) *reinterpret_cast<double2 *>(wrow + 8 * I) = make_double2(acc[t][0], acc[t][1]);
          }
        }
      }
      __syncthreads();
    }
    WW_STAMP();   // +4: products written
    WW_STAMP();   // +4: products written
    const long long tp2 = (P.dbg && tid == 0) ? clock64() : 0;
    // ================================================================ the chains run this batch on warps [0, NCW) while the
    // other warps make the draws of the next one
    WwItem nx = it;
    ww_next(nx, P, TC, NB, ngroups, nwork);
    if (warp >= NCW) {
      if (nx.valid && overlap) pre(nx, tid - NCW * 32, n_pre, 2);
    } else
    {
      constexpr int CPW = 32 / LPC;                     // chains per warp
      const int sub = lane / LPC, g = lane - sub * LPC;
      const int cs = warp * CPW + sub;                  // chain slot in the CTA
      const bool cwarp = warp * CPW < nch_cta;
      if (cwarp) {
        const bool valid = cs < nch_cta;
        const int csv = valid ? cs : nch_cta - 1;       // lanes of a missing chain shadow the last one (no stores)
        const int c_local = cta_chain0 + csv;
        const uint32_t c_global = (uint32_t)(P.cfg.chain_begin + c_local);
        const int i0 = 4 * g;
        const bool own = i0 < ld;
        double x0[4] = {0, 0, 0, 0}, u0[4] = {0, 0, 0, 0};
        double last_prior = cst[csv * 4 + 1], last_like = cst[csv * 4 + 2];
        if (own) {
          const double2 a = *reinterpret_cast<const double2 *>(Xs + csv * ld + i0), b = *reinterpret_cast<const double2 *>(Xs + csv * ld + i0 + 2);
          x0[0] = a.x; x0[1] = a.y; x0[2] = b.x; x0[3] = b.y;
          const double *us = do_refresh ? Wc + (size_t)(ncol + csv) * ld + i0 : Us + csv * ld + i0;
          const double2 c = *reinterpret_cast<const double2 *>(us), e = *reinterpret_cast<const double2 *>(us + 2);
          u0[0] = c.x; u0[1] = c.y; u0[2] = e.x; u0[3] = e.y;
        }
        double ntn_last = nan_to_num(1.0 * last_like + last_prior);
        const double zeta = P.cfg.zeta;
        const double a2sel = lsum_mma_sel<LPC>(lane);
        double *trow_ptr = P.tr.trace + ((size_t)c_local * P.tr.trace_iters + (trace_row0 + done)) * ld + i0;
        double *lrow_ptr = P.tr.trace_logp + (size_t)c_local * P.tr.trace_iters + (trace_row0 + done);
        uint32_t *drow_ptr = P.tr.decisions ? P.tr.decisions + (size_t)c_local * P.tr.trace_iters + (trace_row0 + done) : nullptr;
#pragma unroll 1
        for (int itb = 0; itb < nb; ++itb) {
          const int col = csv * nb + itb;
          const uint32_t mt = meta[col];
          const double lu = logu[col];
          const int run_snooker = (mt >> 8) & 1;
          const double *js = Jc + (size_t)col * ld + i0, *ws = Wc + (size_t)col * ld + i0;
          double prop[4] = {0, 0, 0, 0}, un[4] = {0, 0, 0, 0}, snk_logp = 0.0, cur = 0.0;
          if (own) {
            // prop = q0 + e*gamma*diff + zeta (Dream.py:717); Q(prop) = |u + L^T dx|^2
            const double2 j01 = *reinterpret_cast<const double2 *>(js), j23 = *reinterpret_cast<const double2 *>(js + 2);
            const double2 w01 = *reinterpret_cast<const double2 *>(ws), w23 = *reinterpret_cast<const double2 *>(ws + 2);
            const float4 nn = *reinterpret_cast<const float4 *>(Nz + (size_t)col * ld + i0);
            // (zeta = 0.0 + zeta * n: the stored normals carry no -0, so the product needs no `0.0 +`)
            prop[0] = (x0[0] + j01.x) + zeta * (double)nn.x;
            prop[1] = (x0[1] + j01.y) + zeta * (double)nn.y;
            prop[2] = (x0[2] + j23.x) + zeta * (double)nn.z;
            prop[3] = (x0[3] + j23.y) + zeta * (double)nn.w;
            un[0] = u0[0] + w01.x; un[1] = u0[1] + w01.y; un[2] = u0[2] + w23.x; un[3] = u0[3] + w23.y;
          }
          double part = 0.0;
          if (__any_sync(0xffffffffu, run_snooker)) {
            // snooker_update, Dream.py:827-835 (single-point form); J slot = z, W slot = L^T z, z1 - z2 read from the archive
            double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
            if (run_snooker && own) {
              const double2 j01 = *reinterpret_cast<const double2 *>(js), j23 = *reinterpret_cast<const double2 *>(js + 2);
              a[0] = j01.x; a[1] = j01.y; a[2] = j23.x; a[3] = j23.y;
              const int slot = dpr[col];
              if (slot >= 0) {      // z1 - z2 staged by V2
                const double2 b01 = *reinterpret_cast<const double2 *>(pool + (size_t)slot * ld + i0);
                const double2 b23 = *reinterpret_cast<const double2 *>(pool + (size_t)slot * ld + i0 + 2);
                b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
              } else {              // pool full: read the two rows here
                const double *z1 = zrow(rows[3 * col + 1]) + i0, *z2 = zrow(rows[3 * col + 2]) + i0;
                const double2 p01 = __ldcg(reinterpret_cast<const double2 *>(z1)), p23 = __ldcg(reinterpret_cast<const double2 *>(z1) + 1);
                const double2 q01 = __ldcg(reinterpret_cast<const double2 *>(z2)), q23 = __ldcg(reinterpret_cast<const double2 *>(z2) + 1);
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
            D = lsum_mma<LPC>(D, a2sel);
            S = lsum_mma<LPC>(S, a2sel);
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
            // |prop - z|^2 and Q' together
            nn = lsum_mma<LPC>(nn, a2sel);
            part = lsum_mma<LPC>(part, a2sel);
            {
              // log |prop - z| and log |x - z| (Dream.py:326-332, 834-835) in ONE sqrt + log pass: even lanes take the first,
              // odd lanes the second (a warp instruction costs the same for one lane as for 32), then neighbours swap
              const double nrm = sqrt((lane & 1) ? D : nn);
              const double lg = (nrm != 0 ? log(nrm) : 0.0) * (d - 1);
              const double lo = __shfl_xor_sync(0xffffffffu, lg, 1);
              if (run_snooker) { snk_logp = (lane & 1) ? lo : lg; cur = (lane & 1) ? lg : lo; }
            }
          } else {
            part = fma(un[1], un[1], un[0] * un[0]) + fma(un[3], un[3], un[2] * un[2]);
            part = lsum_mma<LPC>(part, a2sel);
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
          const bool appending = w_append && done + itb == wn - 1;
          if (own && valid) {
            *reinterpret_cast<double2 *>(trow_ptr) = make_double2(x0[0], x0[1]);
            *reinterpret_cast<double2 *>(trow_ptr + 2) = make_double2(x0[2], x0[3]);
            if (appending) {   // record_history: the last iteration of the launch
              double *zr = P.st.Z + (size_t)(M + c_global) * ld + i0;
              *reinterpret_cast<double2 *>(zr) = make_double2(x0[0], x0[1]);
              *reinterpret_cast<double2 *>(zr + 2) = make_double2(x0[2], x0[3]);
              // replicas over NVLink: the chain pushes its own row (its last iteration of the window is over: the
              // system-scope fence below delays nothing but the slowest chain of the group)
                for (int pz = 0; pz < P.npeers; ++pz) {
                  double *zp = P.peer_Z[pz] + (size_t)(M + c_global) * ld + i0;
                  *reinterpret_cast<double2 *>(zp) = make_double2(x0[0], x0[1]);
                  *reinterpret_cast<double2 *>(zp + 2) = make_double2(x0[2], x0[3]);
                }
            }
          }
          if (appending && !counters && P.publish_k) {
            // one launch per window on several GPUs: the chain's warp has stored the replicas itself; the rows are visible
            // system-wide before the chain counts itself.  (Several windows per launch: no fence on the chains' path at all --
            // one thread publishes the group's append after the batch, see below.)
            __threadfence_system();
            __syncwarp();
            if (g == 0 && valid) peer_chain_appended(P.peer_counter, (unsigned)P.cfg.nchains_local, P.peer_flag, P.npeers, P.publish_k);
          }
          if (g == 0 && valid) {
            *lrow_ptr = last_like + last_prior;
            // decision word (dreamzs_common.cuh pack_decision): snooker, CR index, gamma level, one DE pair, gamma == 1
            if (drow_ptr)
              *drow_ptr = (uint32_t)changed | ((mt >> 7) & 2u) | ((mt & 15u) << 2) | (((mt >> 4) & 15u) << 6) | (1u << 10) |
                          (((mt >> 9) & 1u) << 18) | ((uint32_t)accepted << 19);
          }
          trow_ptr += ld; lrow_ptr += 1; if (drow_ptr) drow_ptr += 1;
          WW_STAMP();   // c: iteration done
        }
        // park the chain state for the next batch / the epilogue
        if (own && valid) {
          *reinterpret_cast<double2 *>(Xs + cs * ld + i0) = make_double2(x0[0], x0[1]);
          *reinterpret_cast<double2 *>(Xs + cs * ld + i0 + 2) = make_double2(x0[2], x0[3]);
          *reinterpret_cast<double2 *>(Us + cs * ld + i0) = make_double2(u0[0], u0[1]);
          *reinterpret_cast<double2 *>(Us + cs * ld + i0 + 2) = make_double2(u0[2], u0[3]);
        }
        if (g == 0 && valid) { cst[cs * 4 + 1] = last_prior; cst[cs * 4 + 2] = last_like; }
      }
    }
    WW_STAMP();   // +5: chains advanced
    const long long tp3 = (P.dbg && tid == 0) ? clock64() : 0;
    __syncthreads();   // the slots are free for the next batch; parked states and the next batch's draws are visible
    if (counters && w_append && it.last_batch && tid == WW_THREADS - 1) {
      // Several windows per launch: the group's chains have stored their appended rows (and pushed them to the peers'
      // replicas) before the barrier; ONE thread fences at GPU scope -- measured 17 % of the chain phase when every chain's
      // warp did it on its own last iteration -- and publishes the append: to the local readers (gdone, counters), to the
      // confirmer CTA (counters) and, in this rank's shared header, to the peers that fetch rows from this archive (my_pub).
#ifdef DZ_WW_SC_FENCE
      __threadfence();
#else
      asm volatile("fence.acq_rel.gpu;" ::: "memory");   // (a release is all the publication needs; __threadfence() is fence.sc)
#endif
      if (P.ww_gdone) atomicAdd(P.ww_gdone + it.grp, (uint32_t)nch_cta);
      if (P.my_pub) *reinterpret_cast<volatile uint64_t *>(P.my_pub + it.grp) = P.ww_k0 + (uint64_t)blk + 1u;
      atomicAdd(counters + blk, (uint32_t)nch_cta);
    }
    if (P.dbg && tid == 0) {   // profiling aid: cycles of every CTA's batches by phase, summed over the grid
      unsigned long long *acc = reinterpret_cast<unsigned long long *>(P.dbg) + 60;
      const long long tp4 = clock64();
      atomicAdd(acc + 0, 1ull);                                  // batches
      atomicAdd(acc + 1, (unsigned long long)(tp1 - tp0));       // columns (V2)
      atomicAdd(acc + 2, (unsigned long long)(tp2 - tp1));       // products (M)
      atomicAdd(acc + 3, (unsigned long long)(tp3 - tp2));       // chains of warp 0 (C)
      atomicAdd(acc + 4, (unsigned long long)(tp4 - tp3));       // waiting for the other warps (slower chains / the next batch's draws)
      atomicAdd(acc + 5, (unsigned long long)(tp0 - tprev));     // loop top -> rows requested (TMA issue, states)
      tprev = tp4;
    }
    if (!overlap && nx.valid) {
      pre(nx, tid, WW_THREADS, 1);
      __syncthreads();
    }
    if (it.last_batch && (!resident || it.last_window)) {   // chain states -> global memory
      for (int i = tid; i < nch_cta * nch; i += WW_THREADS) {
        const int cs = fdiv20(i, L.m_nch), q = i - cs * nch;
        const int c_local = cta_chain0 + cs;
        double2 *xr = reinterpret_cast<double2 *>(P.st.X + (size_t)c_local * ld + 4 * q);
        const double2 *xs = reinterpret_cast<const double2 *>(Xs + (size_t)cs * ld + 4 * q);
        xr[0] = xs[0]; xr[1] = xs[1];
        double2 *ur = reinterpret_cast<double2 *>(P.st.gauss_U + (size_t)c_local * ld + 4 * q);
        const double2 *us = reinterpret_cast<const double2 *>(Us + (size_t)cs * ld + 4 * q);
        ur[0] = us[0]; ur[1] = us[1];
      }
      if (tid < nch_cta) {
        P.st.last_prior[cta_chain0 + tid] = cst[tid * 4 + 1];
        P.st.last_like[cta_chain0 + tid] = cst[tid * 4 + 2];
      }
      if (!resident) __syncthreads();   // before the next group's states overwrite the staging area
    }
    it = nx;
  }
#undef WW_STAMP
}

// ---------------------------------------------------------------- host side
// u = L^T x for every local chain from the packed factor (dreamzs_init_logp; a window launch may start at any iteration)
__global__ void __launch_bounds__(128) dreamzs_whiten_kernel(const StepParams P) {
  const int d = P.cfg.ndim, ld = P.cfg.ld, nK = ld / 4;
  const int c = blockIdx.x;
  extern __shared__ double xs_w[];
  for (int j = threadIdx.x; j < ld; j += blockDim.x) xs_w[j] = P.st.X[(size_t)c * ld + j];
  __syncthreads();
  for (int i = threadIdx.x; i < ld; i += blockDim.x) {
    double acc = 0.0;
    if (i < d) {
      const int I = i >> 3;
      const double *tile0 = P.st.gauss_L + (size_t)wwin_tile0(nK, I) * 32;
      for (int j = 8 * I; j < d; ++j) {      // L[j][i] lives in tile (I, k = j/4), lane (j%4) + 4 (i%8)
        const int k = j >> 2;
        const double lji = tile0[(size_t)(k - 2 * I) * 32 + (j & 3) + 4 * (i & 7)];
        acc = fma(lji, xs_w[j], acc);
      }
    }
    P.st.gauss_U[(size_t)c * ld + i] = acc;
  }
}

struct WwinPlan { int tc, nb, nsplit, isplit[WW_MAXSPLIT + 1], lpc; size_t smem; WwinLayout layout; };

inline int wwin_nsplit(int ncols, int nI) {
  const int NT = (ncols + 7) / 8;
  int nsplit = NT > 0 ? WW_WARPS / NT : 0;
  if (nsplit > WW_MAXSPLIT) nsplit = WW_MAXSPLIT;
  if (nsplit > nI) nsplit = nI;
  return nsplit;
}

// chains per CTA, iterations per batch and the split of the i-tiles; tc == 0: the kernel is not usable for this shape.
// `tc_force` / `nb_force` > 0 override the choice (experiments).
inline WwinPlan wwin_plan(const dreamzs_config &cfg, int sms, int niter_max, int tc_force = 0, int nb_force = 0) {
  WwinPlan pl{};
  const int ld = cfg.ld, nch = ld / 4, nI = (ld + 7) / 8, nK = ld / 4;
  if (ld > 128 || (ld & 3) || ld < 8 || cfg.nchains_local < 1) return pl;
  pl.lpc = nch <= 8 ? 8 : nch <= 16 ? 16 : 32;
  const int cpw = 32 / pl.lpc;
  const size_t cap = 227 * 1024;
  int nb = niter_max < 1 ? 1 : niter_max > 16 ? 16 : niter_max;
  if (nb_force > 0) nb = nb_force;
  int tc_want = (cfg.nchains_local + sms - 1) / sms;
  if (tc_force > 0) tc_want = tc_force;
  if (tc_want > WW_WARPS * cpw) tc_want = WW_WARPS * cpw;
  for (;;) {
    // largest tc <= tc_want that fits shared memory and keeps the product units within the CTA's warps
    int tc = tc_want, nsplit = 0;
    for (; tc >= 1; --tc) {
      nsplit = wwin_nsplit(tc * nb + tc, nI);
      if (nsplit >= 1 && nsplit * WW_MAXI >= nI && (size_t)wwin_layout(cfg.ndim, ld, tc, nb, cfg.ngamma).bytes <= cap) break;
    }
    if (tc >= 1) {
      // split the i-tiles into nsplit ranges, each of 1..WW_MAXI tiles, minimising the heaviest range (tile I costs
      // nK - 2 I products): exhaustive over the boundary choices
      auto wsum = [&](int a, int b) { int w = 0; for (int I = a; I < b; ++I) w += (nK - 2 * I > 0 ? nK - 2 * I : 0); return w; };
      int best = 1 << 30, cut[WW_MAXSPLIT + 1];
      cut[0] = 0; cut[nsplit] = nI;
      for (int c1 = 1; c1 <= nI; ++c1)
        for (int c2 = c1; c2 <= nI; ++c2)
          for (int c3 = c2; c3 <= nI; ++c3) {
            const int cs[3] = {c1, c2, c3};
            bool pinned = true;                                          // cuts beyond nsplit-1 are pinned to nI
            for (int q = nsplit - 1; q < 3; ++q) pinned = pinned && cs[q] == nI;
            if (!pinned) continue;
            for (int q = 0; q < nsplit - 1; ++q) cut[q + 1] = cs[q];
            int worst = 0;
            bool fits = true;
            for (int q = 0; q < nsplit; ++q) {
              const int n = cut[q + 1] - cut[q];
              if (n < 1 || n > WW_MAXI) fits = false;
              const int w = wsum(cut[q], cut[q + 1]);
              if (w > worst) worst = w;
            }
            if (fits && worst < best) { best = worst; for (int q = 0; q <= nsplit; ++q) pl.isplit[q] = cut[q]; }
          }
      if (best < (1 << 30)) {
        pl.tc = tc; pl.nb = nb; pl.nsplit = nsplit;
        pl.layout = wwin_layout(cfg.ndim, ld, tc, nb, cfg.ngamma);
        pl.smem = (size_t)pl.layout.bytes;
        return pl;
      }
    }
    if (nb == 1 || nb_force > 0) { pl.tc = 0; return pl; }
    nb = nb > 10 ? 10 : nb > 5 ? 5 : nb - 1;     // fewer iterations per batch
  }
}

template <int LPC>
int launch_wwin_t(StepParams &P, const WwinPlan &pl, int sms, cudaStream_t stream) {
  P.ww_confirm = 0;
  auto kern = dreamzs_wwin_kernel<LPC>;
  static size_t smem_set[64] = {0};
  if (ensure_dynamic_smem(kern, pl.smem, smem_set) != DREAMZS_OK) return DREAMZS_E_LAUNCH;
  const int ngroups = (P.cfg.nchains_local + pl.tc - 1) / pl.tc;
  if (!P.ww_sync) {   // one window: no CTA waits for another, any grid will do
    kern<<<ngroups, WW_THREADS, pl.smem, stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
  }
  // several windows: CTAs wait for rows other CTAs append, so all of them must be resident: at most one wave, a CTA
  // walks the chain groups grid-stride.  A plain launch (a cooperative one would not overlap the copies that drain the
  // previous chunk of samples): should other work hold SMs, the missing CTAs start when it ends; waits time out, not hang.
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WW_THREADS, pl.smem) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    return DREAMZS_E_LAUNCH;
  }
  const int cap = sms * per_sm;
  int grid = ngroups < cap ? ngroups : cap;
  if (P.npeers > 0 && cap >= 2) {   // one more CTA (or the last one) confirms this rank's blocks to the peers
    grid = ngroups + 1 < cap ? ngroups + 1 : cap;
    P.ww_confirm = 1;
  }
  kern<<<grid, WW_THREADS, pl.smem, stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

}  // namespace dreamzs
