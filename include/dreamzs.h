/*
 * dreamzs.h -- C ABI of the B200-native MT-DREAM(ZS) step path (libdreamzs.so).
 *
 * The reference (LoLab-MSM/PyDREAM) is pure Python and has no FFI: the boundary this
 * library sits behind is the Python operator interface
 *     Dream.astep(q0, T, last_loglike, last_logprior) -> (q_new, log_prior, log_like)
 *                                                          pydream/Dream.py:193-422
 *     _sample_dream's per-chain loop over astep             pydream/core.py:89-129
 *     Gelman_Rubin(sampled_parameters) -> Rhat[d]           pydream/convergence.py:3-20
 * Every entry point below names the reference code it replaces.  INTEGRATION.md shows
 * the ctypes stub a PyDREAM maintainer would add to bind them.
 *
 * Conventions
 *   - all pointers marked "device" are CUDA device pointers owned by the caller; the
 *     library never allocates, frees or keeps them; there is no global state (except the
 *     profiling aid dreamzs_debug_set_phase_buffer).
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), does
 *     not synchronise the host, and returns 0 on success or a negative DREAMZS_E_* code.
 *   - all floating-point state is float64; chain ids are GLOBAL ids (shard-independent
 *     random streams: results do not depend on how chains are split over GPUs).
 *   - random numbers: Philox4x32-10, key=(seed lo, seed hi),
 *     counter=(block, call_no<<3|stream, iteration, global chain id); see
 *     DESIGN.md "RNG contract".
 */
#ifndef DREAMZS_H
#define DREAMZS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DREAMZS_ABI_VERSION 4

/* status codes */
#define DREAMZS_OK 0
#define DREAMZS_E_BADARG (-1)     /* inconsistent sizes / null pointers / unsupported option */
#define DREAMZS_E_LAUNCH (-2)     /* cudaGetLastError() != cudaSuccess after the launch */
#define DREAMZS_E_UNSUPPORTED (-3)/* option combination the kernels do not implement */

/* limits baked into the kernels */
#define DREAMZS_MAX_NCR 16
#define DREAMZS_MAX_NGAMMA 8
#define DREAMZS_MAX_DEPAIRS 8
#define DREAMZS_MAX_MULTITRY 16
#define DREAMZS_MAX_NDIM 1024
#define DREAMZS_GAUSS_REFRESH_WINDOWS 4
#define DREAMZS_MAX_PEERS 8
#define DREAMZS_SYNC_GROUP_WORDS 4096

/* dreamzs_config.flags */
#define DREAMZS_FLAG_ALL_FLAT 1  /* every prior is FLAT: the kernels skip prior evaluation and bounds */
#define DREAMZS_FLAG_GENERIC_KERNEL 2 /* always use the generic lane-group kernel (A/B testing of the specialised ones) */
#define DREAMZS_FLAG_NO_WINDOW_KERNEL 4 /* do not use the dense-Gaussian window kernel (A/B testing) */

/* analytic log-likelihoods evaluated in-register (pydream_b200/targets.py) */
enum dreamzs_target_kind {
  DREAMZS_TARGET_CONSTANT = 0,       /* table = [value]                                  */
  DREAMZS_TARGET_GAUSSIAN_DENSE = 1, /* table = [log_F, 0, invC^T (d rows, row stride ld)] */
  DREAMZS_TARGET_MIXTURE = 2,        /* table = [log_F0, log_F1, mu0[d], mu1[d]]         */
  DREAMZS_TARGET_BANANA = 3,         /* table = [b, var1]                                */
  DREAMZS_TARGET_SUMSHIFT = 4,       /* table = [shift]                                  */
  DREAMZS_TARGET_EXTERNAL = 5        /* log-likelihood supplied by the caller (split step) */
};

/* per-dimension prior kinds (pydream/parameters.py:19-70 restricted to closed forms) */
enum dreamzs_prior_kind {
  DREAMZS_PRIOR_FLAT = 0,    /* FlatParam: log prior 0, bounds +-inf            */
  DREAMZS_PRIOR_NORMAL = 1,  /* scipy.stats.norm(loc=a, scale=b)                */
  DREAMZS_PRIOR_UNIFORM = 2  /* scipy.stats.uniform(loc=a, scale=b): [a, a+b]   */
};

/* Options of the sampler: the subset of Dream.__init__ (pydream/Dream.py:63-191) that
 * the step reads, plus the shard geometry. */
typedef struct dreamzs_config {
  int32_t abi_version;      /* DREAMZS_ABI_VERSION */
  int32_t ndim;             /* total_var_dimension, Dream.py:81-83 */
  int32_t ld;               /* row stride (doubles) of Z, X and the trace; multiple of 4, >= ndim */
  int32_t nchains_global;   /* N over all shards */
  int32_t chain_begin;      /* first global chain id of this shard */
  int32_t nchains_local;    /* chains in this shard */
  int32_t nCR;              /* Dream.py:108-112 */
  int32_t ngamma;           /* gamma_levels, Dream.py:120 */
  int32_t nDEpairs;         /* len(self.DEpairs), Dream.py:150 */
  int32_t multitry;         /* 1 = off, Dream.py:155-161 */
  int32_t hardboundaries;   /* Dream.py:80 */
  int32_t history_thin;     /* Dream.py:188 */
  int32_t target_kind;      /* enum dreamzs_target_kind */
  int32_t flags;            /* DREAMZS_FLAG_* */
  double snooker;           /* Dream.py:152 */
  double p_gamma_unity;     /* Dream.py:153 */
  double lamb;              /* Dream.py:164 */
  double zeta;              /* Dream.py:165 */
  uint64_t seed;            /* Philox key */
} dreamzs_config;

/* Device-resident sampler state (all device pointers). */
typedef struct dreamzs_state {
  double *Z;                 /* archive, rows x ld (Dream_shared_vars.history, core.py:281) */
  int64_t Z_capacity_rows;
  double *X;                 /* current positions, nchains_local x ld */
  double *last_prior;        /* nchains_local (Dream.last_prior)  */
  double *last_like;         /* nchains_local (Dream.last_like)   */
  const double *cr_probs;    /* nCR    (Dream_shared_vars.cross_probs)        */
  const double *gamma_probs; /* ngamma (Dream_shared_vars.gamma_level_probs)  */
  const double *gamma_table; /* ngamma x nDEpairs x ndim (Dream.gamma_arr, Dream.py:173-179) */
  const double *target_table;/* see dreamzs_target_kind */
  const int32_t *prior_kind; /* ndim */
  const double *prior_a;     /* ndim */
  const double *prior_b;     /* ndim */
  const double *mins;        /* ndim (Dream.mins, Dream.py:86-105) */
  const double *maxs;        /* ndim */
  /* Optional (may be NULL): carried state of the dense-Gaussian window kernel, maintained by
   * dreamzs_init_logp / dreamzs_step: gauss_Y[c] = invC x_c (nchains_local x ld), gauss_Q[c] = x_c . invC x_c.
   * When both are given (flat priors, one DE pair, no multi-try, 64 < ld <= 128) the quadratic form of a
   * proposal is updated incrementally, Q(x+dx) = Q + 2 dx.y + dx.(invC dx), and re-derived from x every
   * DREAMZS_GAUSS_REFRESH_WINDOWS windows of history_thin iterations. */
  double *gauss_Y;
  double *gauss_Q;
  /* Optional (may be NULL): whitened form of the dense Gaussian, used by the whitened window kernel (flat priors, one
   * DE pair, no multi-try, ld <= 128) in preference to gauss_Y / gauss_Q.  invC = L L^T (L lower triangular);
   * gauss_L is L packed for the fp64 MMA path: for i-tile I = 0 .. ceil(ld/8)-1 and k = 2I .. ld/4-1 one tile of 32
   * doubles, entry (j % 4) + 4 (i % 8) = L[4k + j%4][8I + i%8] (zero outside the matrix), tiles in (I, k) order
   * (dreamzs_whiten_doubles(ld) doubles, 16-byte aligned).  gauss_U[c] = L^T x_c (nchains_local x ld), maintained by
   * dreamzs_init_logp / dreamzs_step: Q(x + dx) = |u + L^T dx|^2, re-derived from x every
   * DREAMZS_GAUSS_REFRESH_WINDOWS windows. */
  const double *gauss_L;
  double *gauss_U;
  /* Optional (may be NULL): scratch words for launches that span several windows (dreamzs_run with the whitened
   * window kernel): 16 + 2 (windows per launch + 2) + DREAMZS_SYNC_GROUP_WORDS uint32.  With it dreamzs_run issues ONE
   * persistent launch for a whole span of iterations; the CTAs synchronise through these words only where a sampled
   * row was appended inside the launch (per window, and -- in the last DREAMZS_SYNC_GROUP_WORDS words -- per group of
   * chains that shares a CTA).  Without it every window is its own launch. */
  uint32_t *sync_ws;
  int64_t sync_ws_words;
  /* Optional (may be NULL): scratch of the two-stage steps, dreamzs_draw_ws_bytes(cfg, iterations per launch) bytes, 16-byte
   * aligned.  With it a window is two (single try) or three (multi-try) kernels: the draw kernel(s) make every draw of
   * the window (decisions, archive gathers, e, zeta, crossover masks, the Metropolis / selection uniforms) for all
   * (chain, iteration[, point]) in parallel and leave them here, the chain kernel walks the chains (proposal assembly,
   * bounds, log-density, selection, acceptance).  Applies to multitry = 1 with any analytic target (a window is cut into
   * sub-spans when the scratch holds fewer iterations) and to multitry = k > 1 when a point fits 8 lanes x 2 chunks
   * (ld <= 64) and k points fit a warp; the dense-Gaussian window kernels do not use it.  Without it the fused kernels run. */
  double *draw_ws;
  int64_t draw_ws_bytes;
} dreamzs_state;

/* Per-launch outputs (device pointers; any may be NULL except trace/trace_logp). */
typedef struct dreamzs_trace {
  double *trace;        /* nchains_local x trace_iters x ld : sampled_params, core.py:99,114 */
  double *trace_logp;   /* nchains_local x trace_iters     : log_ps = like + prior, core.py:115 */
  uint32_t *decisions;  /* nchains_local x trace_iters : bit0 accept (state changed), bit1 snooker,
                           bits2-5 CR index, bits6-9 gamma-level index, bits10-13 DE pairs,
                           bits14-17 selected multi-try index, bit18 gamma==1 flag used by CR adaptation,
                           bit19 metropolis accepted, bit20 (tempered runs, swap rows only) state exchanged
                           with another chain */
  int64_t trace_iters;  /* iterations the trace buffers hold per chain */
  int64_t trace_offset; /* trace row that iteration `iter_begin` writes to */
} dreamzs_trace;

#define DREAMZS_DECISION_SWAPPED (1u << 20)

int dreamzs_abi_version(void);

/* Length (doubles) of the packed whitening factor dreamzs_state.gauss_L for row stride ld. */
int64_t dreamzs_whiten_doubles(int32_t ld);

/* Bytes of dreamzs_state.draw_ws for launches of up to `niter` iterations (0: the two-stage multi-try step does not apply
 * to this configuration). */
int64_t dreamzs_draw_ws_bytes(const dreamzs_config *cfg, int32_t niter);

/* RNG contract, normal variates (stream 2; DESIGN.md "RNG contract"): the 4 * nblocks float32 normals of call
 * `call_no` of chain `chain` at iteration `iter` under `seed`, written to out (device, float32).  Lets a test pin the
 * kernels' Box-Muller to the oracle's bit for bit. */
int dreamzs_rng_normals(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t call_no, int32_t nblocks, float *out,
                        void *stream);

/* Profiling aid, not part of the sampler: CTA 0 of the window kernels writes clock64() stamps of its phases into
 * `device_ptr` (>= 96 int64; NULL switches it off; entries 60.. accumulate per-phase cycles over all CTAs).  Process-wide; the only global state of the library. */
void dreamzs_debug_set_phase_buffer(void *device_ptr);

/* Initial log-prior / log-likelihood of the current positions X (first-call branch of
 * astep, Dream.py:266-268 -> Model.total_logp, pydream/model.py:17-32). */
int dreamzs_init_logp(const dreamzs_config *cfg, const dreamzs_state *st, void *stream);

/* `niter` fused iterations iter_begin .. iter_begin+niter-1 of Dream.astep for every
 * local chain (Dream.py:193-362: set_snooker/set_CR/set_DEpair/set_gamma_level,
 * generate_proposal_points, snooker_update, Model.total_logp on analytic targets,
 * mt_evaluate_logps / mt_choose_proposal_pt, metrop_select, record_history).
 * `archive_rows` = count + nseedchains visible to the FIRST iteration.  The archive must
 * be constant over the launch except for the append made by the LAST iteration of the
 * launch: the caller guarantees that at most the last iteration satisfies
 * iter % history_thin == 0.  That iteration writes chain c's new state to row
 * archive_rows + (global chain id) (record_history, Dream.py:919-938, in chain order). */
int dreamzs_step(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr,
                 int64_t iter_begin, int32_t niter, int64_t archive_rows, void *stream);

/* Parallel tempering (_sample_dream_pt, pydream/core.py:131-236; run_dream(..., tempering=True)).  One iteration
 * of the driver is
 *   dreamzs_step_tempered : Dream.astep(q0, T, last_loglike, last_logprior) (Dream.py:193) for every local chain
 *                           at temperature[c] (nchains_local doubles, device): every log-posterior of the step is
 *                           T * log_like + log_prior (Dream.py:243, 268, 274, 279, 303, 899), last_like / last_prior
 *                           stay untempered (Dream.py:345-347) and trace_logp receives T * like + prior (core.py:176).
 *                           Writes trace row tr->trace_offset; appends when iter % history_thin == 0.
 *   dreamzs_pt_swap       : the proposed exchange of core.py:183-225.  The driver's draws come from the stream of the
 *                           pseudo-chain 0xFFFFFFFF: the pair (first, second) by the random.sample contract
 *                           (np.random.choice(nchains, 2, replace=False), core.py:183), then one 53-bit uniform
 *                           (core.py:195); alpha = (T1 l2 + T2 l1) - (T1 l1 + T2 l2); when log(u) < alpha the two
 *                           chains exchange state, last_like and last_prior.  Every chain is then recorded again in
 *                           trace row tr->trace_offset (row trace_offset - 1 must hold the step's record; trace_logp
 *                           travels with the state, decisions get DREAMZS_DECISION_SWAPPED for the two chains).
 *                           swap_ws: 8 doubles of device scratch, left as (first, second, accepted, alpha, ...).
 *                           All chains must be local (the pair may be any two chains).
 * A tempered run therefore holds 2 trace rows per iteration, as the reference returns them (core.py:145-146). */
int dreamzs_step_tempered(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr,
                          int64_t iter, int64_t archive_rows, const double *temperature, void *stream);
int dreamzs_pt_swap(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr, int64_t iter,
                    const double *temperature, double *swap_ws, void *stream);

/* Split step for likelihoods the caller evaluates (Model.total_logp calls the user's likelihood, pydream/model.py:30;
 * target_kind DREAMZS_TARGET_EXTERNAL).  Without multi-try one iteration of every local chain is
 *   dreamzs_propose : decisions, archive gather, DE / snooker proposal, crossover, boundary handling, log prior
 *                     (everything of Dream.astep up to the likelihood call, Dream.py:246-272); writes
 *                     proposals[nchains_local x ld] and aux[nchains_local x 4] = (log prior of the proposal,
 *                     snooker logp, |x - z|^2, gamma == 1 flag); no state changes
 *   caller          : loglike[c] = log-likelihood of proposals[c] (any device computation on `stream`)
 *   dreamzs_accept  : Metropolis accept, state / trace / decision updates, archive append when
 *                     iter % history_thin == 0 (Dream.py:326-362).
 * With multitry = k > 1 (Dream.py:275-323) a chain has 2k-1 points: proposals[nchains_local x (2k-1) x ld],
 * loglike[nchains_local x (2k-1)], aux[nchains_local x (4k+4)], and the iteration is
 *   dreamzs_propose -> caller fills loglike[:, 0:k] -> [dreamzs_repropose -> caller fills loglike[:, 0:k]]* ->
 *   dreamzs_select (mt_choose_proposal_pt, then the k-1 reference points around the selected proposal go to
 *   proposals[:, k:2k-1]) -> caller fills loglike[:, k:2k-1] -> dreamzs_accept.
 *   dreamzs_repropose is the regenerate loop of Dream.py:278-289: every chain whose k proposals all have a non-finite
 *   log-posterior draws its next batch (the stream goes on where the previous batch left it) into proposals[:, 0:k];
 *   *count (device int32, zeroed by the caller) receives the number of such chains -- repeat until it stays 0 (the
 *   reference has no bound on the loop; the fused kernels stop after 1000 batches).  dreamzs_select sets *error
 *   (device int32) to 1 for a chain that still has no finite proposal.
 * The random stream resumes in each phase where the previous one left it, so the phases consume exactly the draws
 * of the fused step.  dreamzs_init_logp sets last_prior and leaves last_like = 0 for the caller to fill. */
int dreamzs_propose(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows,
                    double *proposals, double *aux, void *stream);
int dreamzs_repropose(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows,
                      double *proposals, double *aux, const double *loglike, int32_t *count, void *stream);
int dreamzs_select(const dreamzs_config *cfg, const dreamzs_state *st, int64_t iter, int64_t archive_rows,
                   double *proposals, double *aux, const double *loglike, int32_t *error, void *stream);
int dreamzs_accept(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr, int64_t iter,
                   int64_t archive_rows, const double *proposals, const double *aux, const double *loglike,
                   void *stream);

/* Replicas of the archive on the other GPUs of the box (pydream/Dream_shared_vars.py `history` is ONE shared
 * array in the reference; here every GPU holds a copy and the copies are kept identical over NVLink).
 * Z[q] / flags[q] are rank q's archive and flag array (DREAMZS_MAX_PEERS uint64, zero-initialised) as mapped
 * into THIS process (cudaIpcOpenMemHandle, see dreamzs_shared_*); Z[rank] must equal dreamzs_state.Z.
 * With peers, the appending iteration of a launch stores every local chain's row into all replicas (plain
 * peer stores from the step kernel, each followed by a system-scope fence); the chain that completes the
 * launch's appends publishes "append #k done" into flags[q][rank] of every peer (st.release.sys), and the
 * next launch that may sample those rows first waits until every flags[rank][q] has reached k
 * (ld.acquire.sys; the wait gives up after DREAMZS_PEER_TIMEOUT_NS and sets *error).  Compute and the
 * exchange are one kernel: no collective call and no extra launch per window. */
typedef struct dreamzs_peers {
  int32_t world, rank;
  double *Z[DREAMZS_MAX_PEERS];
  uint64_t *flags[DREAMZS_MAX_PEERS];
  uint32_t *counter;         /* device uint32 of this rank, zero-initialised: chains that have appended (scratch) */
  int32_t *error;            /* device int32 of this rank, zero-initialised; 1 after a timed-out wait */
  /* 0, or the length of the per-group progress rows that follow the append flags in every rank's block: from
   * flags[q] + DREAMZS_GFLAG_OFFSET on, world rows of gflag_stride uint64 (zero-initialised); rank q uses row q of ITS
   * OWN block: word g = appends the chains of its group g (the chains that share a CTA) have completed.  Launches that
   * span several windows push appended rows to the peers without a system-scope fence on the chains' path; a block is
   * confirmed in flags[q][rank] once per append by one thread.  A reader that samples a peer's row before its block is
   * confirmed polls that group's word in the OWNER's memory and fetches the row from the owner's archive (both over
   * NVLink), so it waits for the chains that write the row and for nothing else. */
  int32_t gflag_stride;
  int32_t reserved;
} dreamzs_peers;
#define DREAMZS_GFLAG_OFFSET 512   /* uint64 words from flags[q] to the per-group flag rows */
#define DREAMZS_PEER_TIMEOUT_NS 10000000000ull

/* Device memory that other processes of the box can map: cudaMalloc + cudaIpcGetMemHandle (zero-filled),
 * cudaIpcOpenMemHandle (peer access enabled lazily), and their inverses.  `handle` is the 64-byte
 * cudaIpcMemHandle_t.  These are the only entry points that allocate; the caller owns the result. */
int dreamzs_shared_alloc(int64_t bytes, void **dev_ptr, void *handle);
int dreamzs_shared_open(const void *handle, void **dev_ptr);
int dreamzs_shared_close(void *dev_ptr);
int dreamzs_shared_free(void *dev_ptr);

/* The chain loop of _sample_dream (pydream/core.py:103-122) for iterations that need no host decision between
 * them: iterations iter_begin .. iter_begin+niter-1 as one dreamzs_step launch per window, a window ending
 * at an iteration t with t % history_thin == 0 (burn-in iterations with adaptation: one launch each, see
 * dreamzs_adapt).  tr->trace_offset is the trace row of
 * iter_begin.  `appends_done` = number of appends made so far (archive_rows = seed rows + appends_done *
 * nchains_global).  Sharded runs keep the replicas of the archive identical either through `peers`
 * (NVLink peer stores, see dreamzs_peers) or, when peers is NULL, through `hook(user, first_row, nrows)`,
 * called on the host after each appending launch (it may enqueue stream-ordered work, e.g. an NCCL all-gather
 * of the other shards' rows; it must return 0).  Both may be NULL when nchains_local == nchains_global.
 * *launches (optional) receives the number of kernel launches, *archive_rows_out (optional) the archive
 * size after the run. */
typedef int (*dreamzs_append_hook)(void *user, int64_t first_row, int64_t nrows);

/* Burn-in adaptation inside the native loop (Dream.py:364-401): while adapt != NULL and t <= crossover_burnin
 * every iteration is its own launch, followed -- for 10 < t < crossover_burnin and once more at
 * t == crossover_burnin -- by the reduction stages dreamzs_adapt_colsum/colsq/jumps/finish below.  All
 * pointers are device buffers of the caller: colsum[d], colsq[d], partial[2 nCR + 2 ngamma], workspace
 * (dreamzs_adapt_workspace_bytes), x_entry[nchains_local x ld] (receives the states at entry), the
 * accumulators ncr_updates/delta_m[nCR], ngamma_updates/delta_m_gamma[ngamma] and the probabilities the
 * step reads (the same buffers as dreamzs_state.cr_probs / gamma_probs).  Needs dreamzs_trace.decisions.
 * Sharded runs pass `reduce(user, buffer, count)`, called after colsum, colsq and jumps to sum the buffer
 * over the ranks (stream-ordered, e.g. an NCCL all-reduce); NULL when all chains are local. */
typedef int (*dreamzs_reduce_hook)(void *user, double *buffer, int64_t count);
typedef struct dreamzs_adapt {
  int32_t adapt_crossover, adapt_gamma;
  int64_t crossover_burnin;
  double *colsum, *colsq, *partial;
  void *workspace;
  double *x_entry;
  double *ncr_updates, *delta_m, *cr_probs;
  double *ngamma_updates, *delta_m_gamma, *gamma_probs;
  dreamzs_reduce_hook reduce;
  void *user;
} dreamzs_adapt;

int dreamzs_run(const dreamzs_config *cfg, const dreamzs_state *st, const dreamzs_trace *tr,
                int64_t iter_begin, int64_t niter, int64_t archive_rows, int64_t appends_done,
                const dreamzs_peers *peers, dreamzs_append_hook hook, void *user, const dreamzs_adapt *adapt,
                void *stream, int64_t *launches, int64_t *archive_rows_out);

/* sampled_params / log_ps leave the device (pydream/core.py:81-86 returns them to the caller): stream-ordered
 * copy of `height` rows of `width_bytes` from a pitched device block to a pitched (pinned) host block. */
int dreamzs_copy_d2h_2d(void *dst_host, int64_t dst_pitch_bytes, const void *src_device, int64_t src_pitch_bytes,
                        int64_t width_bytes, int64_t height, void *stream);

/* Crossover-probability (and gamma-level) adaptation for ONE iteration of the burn-in
 * (estimate_crossover_probabilities, Dream.py:451-499; estimate_gamma_level_probs,
 * Dream.py:501-540; set_current_position_arr, Dream.py:424-449), as stream-ordered reduction
 * stages so that chains may be sharded (the caller all-reduces between stages when sharded):
 *   colsum : colsum[d]  = sum over LOCAL chains of X_new[c]
 *   colsq  : colsq[d]   = sum over LOCAL chains of (X_new[c] - colsum_global/N)^2   (np.std is two-pass)
 *   jumps  : partial[2*nCR + 2*ngamma] = for the LOCAL chains, per CR index the number of updates and
 *            the summed squared normalised jump (then the same per gamma level); sd = sqrt(colsq_global/N).
 *            Which chains count follows Dream.py:371-383 from this iteration's decision words
 *            (stride dec_stride); final_update != 0 is the unconditional update at
 *            iter == crossover_burnin (Dream.py:391-401).  x_old: states before the iteration,
 *            chain stride ld_old doubles.
 *   finish : folds the (all-reduced) partials into ncr_updates/delta_m (and the gamma twins) and
 *            renormalises the probabilities once every delta_m is non-zero (Dream.py:483-495).
 * workspace: dreamzs_adapt_workspace_bytes(cfg) bytes of device memory. */
int64_t dreamzs_adapt_workspace_bytes(const dreamzs_config *cfg);
int dreamzs_adapt_colsum(const dreamzs_config *cfg, const double *X_new, double *colsum, void *workspace,
                         void *stream);
int dreamzs_adapt_colsq(const dreamzs_config *cfg, const double *X_new, const double *colsum, double *colsq,
                        void *workspace, void *stream);
int dreamzs_adapt_jumps(const dreamzs_config *cfg, const double *X_new, const double *x_old, int64_t ld_old,
                        const uint32_t *decisions, int64_t dec_stride, const double *colsq,
                        int32_t final_update, int32_t adapt_crossover, int32_t adapt_gamma, double *partial,
                        void *workspace, void *stream);
int dreamzs_adapt_finish(const dreamzs_config *cfg, const double *partial, int32_t adapt_crossover,
                         int32_t adapt_gamma, double *ncr_updates, double *delta_m, double *cr_probs,
                         double *ngamma_updates, double *delta_m_gamma, double *gamma_probs, void *stream);

/* Gelman-Rubin diagnostic (pydream/convergence.py:3-20) in three stream-ordered stages so
 * that chains may be sharded:
 *   chain_stats: per local chain, mean[d] and var[d] (ddof 0) of trace rows nburnin..nsamples-1
 *   caller all-gathers / concatenates the per-chain stats over shards
 *   finish: W = mean_c var, B = var_c(mean) (ddof 0), Rhat = sqrt((W (1-1/nsamples) + B)/W) */
int dreamzs_gr_chain_stats(const double *trace, int64_t nchains, int64_t nsamples, int64_t nburnin,
                           int32_t ndim, int64_t ld, double *chain_mean, double *chain_var,
                           void *stream);
int dreamzs_gr_finish(const double *chain_mean, const double *chain_var, int64_t nchains,
                      int64_t nsamples, int32_t ndim, double *rhat, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DREAMZS_H */
