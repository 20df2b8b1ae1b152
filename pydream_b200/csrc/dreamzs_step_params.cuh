#pragma once
#include <cuda_runtime.h>
#include "../../include/dreamzs.h"

namespace dreamzs {

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
  long long *dbg;         // optional phase-timestamp buffer (dreamzs_debug_set_phase_buffer; profiling aid)
  int32_t gw_append;      // window kernel: the last iteration of the launch appends to the archive
  int32_t gw_refresh;     // window kernel: re-derive gauss_Y / gauss_Q from X at the start of the launch
};

}  // namespace dreamzs
