#pragma once
#include <cuda_runtime.h>
#include "../../include/dreamzs.h"

namespace dreamzs {

// shared-memory carve-up of the whitened window kernel (byte offsets; dreamzs_wwin_kernel.cuh wwin_layout)
struct WwinLayout {
  int32_t nch, nI, nK, ntilesL, ncolmax;
  int32_t oL, oW, oJ, oN, oXs, oUs, oGam, oScr, oLogu, oGsn, oRows, oMbar, oMeta, oDpr, oMask, oUses, oProbs, oCst, oPool, npool, bytes;
  uint32_t m_nch;   // ceil(2^20 / nch): division-free split of a task number
};

struct StepParams {
  dreamzs_config cfg;
  dreamzs_state st;
  dreamzs_trace tr;
  int64_t iter_begin;
  int32_t niter;
  int64_t archive_rows;
  int32_t all_flat;       // every prior is FLAT (skip prior evaluation and bounds)
  int32_t table_in_smem;  // target table staged in shared memory
  int32_t table_doubles;  // length of the target table
  int32_t nslots;         // proposal slots per chain in shared memory
  int32_t init_only;      // dreamzs_init_logp: evaluate logp(X) and return
  int32_t gw_nb;          // window kernel: iterations per batch
  int32_t npeers;         // other GPUs holding a replica of the archive (NVLink peer mappings)
  double *peer_Z[DREAMZS_MAX_PEERS];   // their Z, as mapped in this process
  uint64_t *peer_flag[DREAMZS_MAX_PEERS];   // their flag word for this rank (flags[q] + rank)
  const uint64_t *my_flags;   // this rank's flag array (one word per rank)
  int32_t my_rank, world;
  uint64_t wait_k;        // != 0: before touching the archive wait until every peer has published append #wait_k
  uint64_t publish_k;     // != 0: this launch appends; when all local chains have, publish append #publish_k
  unsigned int *peer_counter;   // chains of this launch that have appended (scratch, returns to 0)
  int32_t *peer_error;
  int32_t ext_phase;      // split step for caller-evaluated likelihoods: 0 fused; single try: 1 propose, 2 accept; multi-try: 1 propose, 2 select, 3 accept
  double *ext_prop;       // [nchains_local x ld] proposals (written by phase 1, read by phase 2)
  double *ext_aux;        // [nchains_local x 4]  log prior, snooker logp, |x - z|^2, gamma == 1 flag
  const double *ext_like; // [nchains_local]      caller's log-likelihood of the proposals (phase 2)
  int32_t *ext_error;     // set to 1 when a multi-try batch has no finite log-posterior (the regenerate loop is not split)
  long long *dbg;         // optional phase-timestamp buffer (dreamzs_debug_set_phase_buffer; profiling aid)
  int32_t gw_append;      // window kernel: the last iteration of the launch appends to the archive
  int32_t gw_refresh;     // window kernel: re-derive gauss_Y / gauss_Q from X at the start of the launch
  WwinLayout ww_L;
  int32_t ww_tc, ww_nb, ww_nsplit, ww_isplit[5];
  uint32_t *ww_sync;      // != NULL: the launch spans several windows; word 0 abort flag, words 16.. chains that made append #j
  uint32_t *ww_gdone;     // != NULL: per group of ww_tc chains, the appends its chains have made in this launch (summed over the chains)
  // per-group progress words (dreamzs_peers.gflag_stride): my_pub[g] = appends group g of THIS rank has completed since the
  // start of the run (written here, in this rank's own memory); peer_pub[pz] = the same words of peer pz, read over NVLink
  uint64_t *my_pub;
  const uint64_t *peer_pub[DREAMZS_MAX_PEERS];
  int32_t gflag_stride;
  int32_t ww_confirm;     // the launch's last CTA confirms this rank's appended blocks to the peers (dreamzs_wwin_kernel.cuh)
  int32_t ww_wcap;        // windows the scratch words hold: counters[ww_wcap] chains appended, then [ww_wcap] chains forwarded to the peers
  uint64_t ww_k0;         // appends made before this launch (peer flags count appends from the start of the run)   // whitened window kernel: chains per CTA, iterations per batch, i-tile ranges of the products
  const double *temperature;   // [nchains_local] per-chain temperature T of astep(q0, T, ...) (Dream.py:193); NULL = 1
};

// point-parallel multi-try kernels (dreamzs_mtp_kernel.cuh), host side: the layout of a configuration -- lanes per point G, chunk rounds R (G R >= ld / 4, R <= 2), as many chains per
// warp as fit
inline bool mtp_layout(int ld, int k, int &G, int &R) {
  const int chunks = ld / 4;
  int best = 0;
  for (int g = 2; g <= 8; g <<= 1)
    for (int r = 1; r <= 2; ++r) {
      if (g * r < chunks || k * g > 32) continue;
      const int cpw = 32 / (k * g);
      const int score = cpw * 4 - r;          // more chains per warp first, then fewer rounds
      if (score > best) { best = score; G = g; R = r; }
    }
  return best > 0;
}

}  // namespace dreamzs
