// Whitened window kernel (dreamzs_wwin_kernel.cuh): instantiations per lanes-per-chain, launch plan, whitening refresh.
#include <stdlib.h>
#include "dreamzs_wwin_kernel.cuh"

using namespace dreamzs;

namespace {
struct PlanKey { int ndim, ld, nchains_local, ngamma, sms, niter; };
struct PlanSlot { PlanKey key; WwinPlan plan; bool used; };

int env_int(const char *name) {
  const char *v = getenv(name);
  return v ? atoi(v) : 0;
}
}  // namespace

// the plan of a launch depends on a handful of integers: a small per-thread cache keeps the search off the launch path
static const WwinPlan &cached_plan(const dreamzs_config &cfg, int sms, int niter) {
  thread_local PlanSlot slots[8] = {};
  thread_local int next = 0;
  const PlanKey key = {cfg.ndim, cfg.ld, cfg.nchains_local, cfg.ngamma, sms, niter};
  for (auto &s : slots)
    if (s.used && s.key.ndim == key.ndim && s.key.ld == key.ld && s.key.nchains_local == key.nchains_local &&
        s.key.ngamma == key.ngamma && s.key.sms == key.sms && s.key.niter == key.niter)
      return s.plan;
  PlanSlot &s = slots[next];
  next = (next + 1) & 7;
  s.key = key; s.used = true;
  // DREAMZS_WW_TC / DREAMZS_WW_NB: experiments only (chains per CTA, iterations per batch)
  s.plan = wwin_plan(cfg, sms, niter, env_int("DREAMZS_WW_TC"), env_int("DREAMZS_WW_NB"));
  return s.plan;
}

int dreamzs_wwin_usable(const dreamzs_config &cfg, int sms, int niter) { return cached_plan(cfg, sms, niter).tc > 0; }

int dreamzs_launch_wwin(StepParams &P, int sms, cudaStream_t stream) {
  const int wmax = P.niter < P.cfg.history_thin ? P.niter : P.cfg.history_thin;     // longest window of the launch
  const WwinPlan &pl = cached_plan(P.cfg, sms, wmax);
  if (pl.tc == 0) return DREAMZS_E_UNSUPPORTED;
  P.ww_tc = pl.tc; P.ww_nb = pl.nb; P.ww_nsplit = pl.nsplit; P.ww_L = pl.layout;
  for (int q = 0; q <= WW_MAXSPLIT; ++q) P.ww_isplit[q] = pl.isplit[q];
  // several windows in one launch: per-group completion counters in the last words of the scratch, per-group flags of the
  // peers -- when the groups fit
  const int ngroups = (P.cfg.nchains_local + pl.tc - 1) / pl.tc;
  P.ww_gdone = (P.ww_sync && ngroups <= DREAMZS_SYNC_GROUP_WORDS) ? P.ww_sync + (P.st.sync_ws_words - DREAMZS_SYNC_GROUP_WORDS) : nullptr;
  if (!P.ww_sync || ngroups > P.gflag_stride) { P.my_pub = nullptr; P.gflag_stride = 0; }
  if (pl.lpc == 32) return launch_wwin_t<32>(P, pl, sms, stream);
  if (pl.lpc == 16) return launch_wwin_t<16>(P, pl, sms, stream);
  return launch_wwin_t<8>(P, pl, sms, stream);
}

// gauss_U = L^T x for every local chain (dreamzs_init_logp)
int dreamzs_launch_whiten(const StepParams &P, cudaStream_t stream) {
  dreamzs_whiten_kernel<<<P.cfg.nchains_local, 128, (size_t)P.cfg.ld * sizeof(double), stream>>>(P);
  return cudaGetLastError() == cudaSuccess ? DREAMZS_OK : DREAMZS_E_LAUNCH;
}

namespace dreamzs { int wwin_ntiles_host(int ld) { return wwin_ntiles(ld); } }
