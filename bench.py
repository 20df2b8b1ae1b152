#!/usr/bin/env python
"""Benchmark of the MT-DREAM(ZS) step path on B200 (BASELINE.json metric: chain-steps/sec).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host threads
    python bench.py --workload c4|c5 --gpus N ...            # the strong-sharded BASELINE configs 4 / 5 as the main line

Main workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): 100-D correlated Gaussian
(dream_ex_ndim_gaussian.py covariance), 1024 chains per GPU, FlatParam prior, reference default options
(snooker .1, DEpairs 1, nCR 3, history_thin 10, multitry off), crossover adaptation off so that the timed
region is the steady-state step.

ONE BENCH STEP = `iters_per_step` sampler iterations of every chain (C2: 1000 iterations = 100 history_thin windows
= 1.024 M chain-steps per GPU; stated in config.workload and applied identically to the CPU arm).  The K-step region is
timed `repeats` times (archive rewound to its seed rows in between, outside the timed region, so that memory stays
bounded) until >= 1 s of steady state has been timed; `value` is computed from the MEDIAN region.  The archive is
pre-seeded with 262144 rows (210 MB > the 126 MB L2) so every timed gather works on an input larger than L2.  Scaling
over GPUs is weak for C2 (1024 chains per GPU, archive replicated, appended rows reach the other replicas as NVLink
peer stores from inside the step kernel; NCCL all-gather when peer mappings are unavailable).

`value`   : K steps through DreamEngine.run with everything resident in HBM, CUDA events, max over ranks, median region.
`e2e`     : the same metric through pydream_b200.core.run_dream with host (numpy) inputs and outputs: archive
            seed + starts uploaded, every sample and log-posterior copied back, inside the timed region.
`roofline`: algorithmic HBM bytes per chain-step x chain-steps per launch / launch duration,
            against the measured copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` (N=1): oracle/dreamzs_oracle.c (a C port of the reference's step) on the host threads, bounded sample;
            `cpu_baseline.reference_python`: the UNMODIFIED reference's own multiprocessing run_dream on the same host
            (baseline/_ref or /root/reference; tools/ref_python_bench.py), N in {3, 8, ncores}, with and without its sleeps.
`other_configs` (N=1): BASELINE configs 3, 4, 5 -- value, roofline, dominant kernel, launch duration -- each on an
            archive larger than L2.
`check`   : at N > 1 a short sharded run is compared bit for bit with the identical single-GPU run on rank 0
            (outside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0
COMMON = dict(snooker=.1, history_thin=10, DEpairs=1, nCR=3, p_gamma_unity=.2, lamb=.05, zeta=1e-12)
# chains: per GPU when weak, in total when strong (BASELINE configs 4 / 5 shard a fixed population)
WORKLOADS = {
    'c2': dict(label='C2: 100-D correlated Gaussian (dense precision), 1024 chains per GPU, fused DE/snooker proposal '
                     '+ logp + accept kernel', d=100, chains=1024, scaling='weak', target='gaussian', nseed=262144,
               iters_per_step=1000, opts=dict(COMMON, multitry=1), seed_dist='box',
               kernel='dreamzs_wwin_kernel<32>'),
    'c3': dict(label='C3: 10-D bimodal Gaussian mixture, 4096 chains per GPU, multi-try 5 + snooker', d=10, chains=4096,
               scaling='weak', target='mixture', nseed=2 ** 21, iters_per_step=100, opts=dict(COMMON, multitry=5),
               seed_dist='normal', kernel='dreamzs_mtchain_kernel<2,2> (after dreamzs_mtdraw_scalars_kernel + dreamzs_mtdraw_kernel<4>, three launches per window)'),
    'c4': dict(label='C4: 200-D twisted Gaussian (banana, b=0.1), 8192 chains in total', d=200, chains=8192,
               scaling='strong', target='banana', nseed=131072, iters_per_step=100, opts=dict(COMMON, multitry=1),
               seed_dist='banana', kernel='dreamzs_stdraw_kernel<32,2> + dreamzs_stchain_kernel<32,2> (two launches per window)'),
    'c5': dict(label='C5: 50-D correlated Gaussian, 65536 chains in total, steady-state step (crossover adaptation during '
                     'burn-in and Gelman-Rubin timed separately: burnin / rhat)', d=50, chains=65536, scaling='strong',
               target='gaussian', nseed=524288, iters_per_step=20, opts=dict(COMMON, multitry=1), seed_dist='box',
               kernel='dreamzs_wwin_kernel<16>'),
}
LOGP_TOL = '|dlogp| <= 1e-12 * max(1, |logp|) (relative reading of the north_star 1e-12: ulp(1e4) is 1.8e-12); decisions bit-exact'


def b_step(d, k, s, delta, thin):
    """Algorithmic HBM bytes per chain-step: state read+write, R gathered archive rows,
    1/thin appended rows, logp read+write."""
    R = (2 * k - 1) * ((1 - s) * 2 * delta + s * 3)
    return 8 * d * (2 + R + 1.0 / thin) + 16


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def make_target(kind, d):
    from pydream_b200 import targets
    if kind == 'gaussian':
        return targets.CorrelatedGaussian.benchmark(d)
    if kind == 'mixture':
        return targets.BimodalMixture.benchmark(d)
    return targets.Banana(d, 0.1)


def synthetic_inputs(wl, nchains, rng_seed=1234, nseed=None):
    """Archive seed: C2 / C5 ~ U(-5,15)^d (the example's Latin-hypercube box, dream_ex_ndim_gaussian.py:17-26,45), C3 ~ N(0,I)
    (mixturemodel.py:30-33), C4 ~ N(0, diag(100,1,..)); chain c starts at seed row c (dream_ex_ndim_gaussian.py:54)."""
    rng = np.random.default_rng(rng_seed)
    d, n = wl['d'], int(nseed or wl['nseed'])
    if wl['seed_dist'] == 'box':
        hist = rng.uniform(-5, 15, size=(n, d))
    else:
        hist = rng.standard_normal(size=(n, d))
        if wl['seed_dist'] == 'banana':
            hist[:, 0] *= 10.0
    return hist, hist[:nchains].copy()


def total_chains(wl, world):
    return wl['chains'] * world if wl['scaling'] == 'weak' else wl['chains']


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        fd, self.path = tempfile.mkstemp(suffix='.csv')
        os.close(fd)
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(',')]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        os.unlink(self.path)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU arms
def host_threads():
    """Threads the CPU arm may use: the affinity mask, clipped by a cgroup CPU quota if one is set."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, per = f.read().split()
            if q != 'max':
                n = min(n, max(1, int(float(q) / float(per))))
    except (OSError, ValueError):
        try:
            with open('/sys/fs/cgroup/cpu/cpu.cfs_quota_us') as f, open('/sys/fs/cgroup/cpu/cpu.cfs_period_us') as g:
                q, per = int(f.read()), int(g.read())
                if q > 0:
                    n = min(n, max(1, q // per))
        except (OSError, ValueError):
            pass
    return max(1, n)


class OracleArm:
    """The CPU arm: oracle/dreamzs_oracle.c (a port of the reference's step path; lock-step sweep split over host
    threads) on the workload `wl` with `nchains` chains."""

    def __init__(self, wl, nchains, nthreads, reserve_iters):
        from oracle import c_oracle
        tgt = make_target(wl['target'], wl['d'])
        hist, starts = synthetic_inputs(wl, nchains)
        thin = wl['opts']['history_thin']
        self.nchains, self.nthreads = nchains, nthreads
        self.s = c_oracle.OracleSampler(wl['d'], nchains, hist, starts, tgt.kind, tgt.table(), seed=SEED, nthreads=nthreads,
                                        capacity_rows=hist.shape[0] + (reserve_iters // thin + 2) * nchains, **wl['opts'])

    def run(self, niter):
        t0 = time.perf_counter()
        self.s.run(niter)
        return time.perf_counter() - t0


def pick_threads(wl, nchains):
    """Try the full thread count and a few smaller ones on a short sample (oversubscribed or throttled hosts run slower
    with more threads) -> (best thread count, its chain-steps/s)."""
    nmax = host_threads()
    cands = sorted({nmax, max(1, nmax // 2), min(nmax, 32), min(nmax, 16), min(nmax, 8)}, reverse=True)
    best, best_rate = cands[-1], 0.0
    for n in cands:
        arm = OracleArm(wl, nchains, n, 64)
        arm.run(2)
        its = max(4, min(30, int(2e5 // nchains)))
        rate = nchains * its / arm.run(its)
        if rate > best_rate:
            best, best_rate = n, rate
    return best, best_rate


def oracle_sample(wl, nchains, budget_s, warmup_iters=3):
    """Bounded sample of the workload on the CPU port: (chain-steps/s, seconds, iterations, threads)."""
    nthreads, rate = pick_threads(wl, nchains)
    iters = int(max(10, min(wl['iters_per_step'] * 4, rate * budget_s / nchains)))
    arm = OracleArm(wl, nchains, nthreads, iters + warmup_iters)
    arm.run(warmup_iters)
    dt = arm.run(iters)
    return nchains * iters / dt, dt, iters, nthreads


def bench_reference(args):
    """The CPU arm on the main arm's config, metric and step definition.  K timed steps after W warm-up steps; a step is a
    BOUNDED SAMPLE of the `iters_per_step` iterations (all of them when the budget allows), sized from a short calibration
    so that the whole run ends within a few minutes; ms_per_step is scaled to the full step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    nchains = total_chains(wl, args.gpus)
    nthreads, rate = pick_threads(wl, nchains)
    ips = wl['iters_per_step']
    sample = int(max(5, min(ips, rate * args.cpu_budget / (nchains * (K + W)))))
    arm = OracleArm(wl, nchains, nthreads, (K + W) * sample)
    for _ in range(W):
        arm.run(sample)
    dts = [arm.run(sample) for _ in range(K)]
    dt = float(np.sum(dts))
    value = nchains * sample * K / dt
    desc = ('each step = %d of the %d iterations of a bench step (%d chains), oracle/dreamzs_oracle.c on %d host threads; '
            '%d warm-up + %d timed steps, %.1f s timed' % (sample, ips, nchains, nthreads, W, K, dt))
    line = dict(impl='reference', metric='chain-steps/sec', value=value, unit='chain-steps/s', n_gpus=args.gpus,
                steps=K, warmup=W, ms_per_step=1e3 * (dt / K) * (ips / sample), higher_is_better=True, scaling=wl['scaling'],
                vs_baseline=None, dtype='f64', data='synthetic',
                config=config_of(wl, nchains, args.gpus, cpu_arm=True),
                cpu_baseline=dict(value=value, unit='chain-steps/s', cores=nthreads, kind='port', sample=desc),
                e2e=dict(value=value, unit='chain-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                sampled_iterations_per_step=sample, gpu_launches=0)
    print(json.dumps(line), flush=True)


def config_of(wl, nchains, world, cpu_arm=False, **extra):
    cfg = dict(workload='%s; one bench step = %d sampler iterations of all chains (%d chain-steps)%s'
                        % (wl['label'], wl['iters_per_step'], wl['iters_per_step'] * nchains,
                           ' -- CPU arm: C port of pydream Dream.astep, lock-step sweep split over host threads' if cpu_arm else ''),
               ndim=wl['d'], nchains=nchains, chains_per_gpu=nchains // world, archive_seed_rows=wl['nseed'],
               iters_per_step=wl['iters_per_step'], chain_steps_per_step=wl['iters_per_step'] * nchains)
    cfg.update(wl['opts'])
    cfg.update(extra)
    return cfg


def reference_python(ncores):
    """The unmodified reference's own multiprocessing run_dream on this host (SURVEY.md 8(d) "CPU baseline timing"):
    sleeps excluded (crossover_burnin > niter) at N in {3, 8, ncores}; the default settings (>= 40 s of sleeps,
    Dream.py:403-407) at N = 3 are started in the background by the caller (see start_reference_python_default)."""
    tool = os.path.join(ROOT, 'tools', 'ref_python_bench.py')
    out = []
    for n in sorted({3, 8, max(3, min(int(ncores), 64))}):
        niter = int(max(100, min(2000, 24000 // n)))
        try:
            p = subprocess.run([sys.executable, tool, '--nchains', str(n), '--niter', str(niter), '--mode', 'nosleep'],
                               capture_output=True, text=True, timeout=240, cwd=tempfile.gettempdir())
            j = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception as e:      # noqa: BLE001
            j = dict(unavailable='%s: %s' % (type(e).__name__, str(e)[:200]))
        j['requested_nchains'] = n
        out.append(j)
    return out


def start_reference_python_default():
    tool = os.path.join(ROOT, 'tools', 'ref_python_bench.py')
    try:
        return subprocess.Popen([sys.executable, tool, '--nchains', '3', '--niter', '1000', '--mode', 'default'],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd=tempfile.gettempdir())
    except OSError:
        return None


# ------------------------------------------------------------------------------------------------ GPU arm
def timed_regions(eng, wl, K, W, barrier, reduce_max, min_total_s=1.0, max_repeats=25):
    """W warm-up steps, then the K-step region repeated until >= min_total_s have been timed (at least 3 regions).
    Returns (list of region ms, launches per region).  The archive is rewound between regions (outside the timing)."""
    import torch
    ips = wl['iters_per_step']
    trace = torch.empty((eng.Nl, ips, eng.ld), dtype=torch.float64, device=eng.device)
    logp = torch.empty((eng.Nl, ips), dtype=torch.float64, device=eng.device)
    for _ in range(W):
        eng.run(ips, trace=trace, logp=logp)
    ms_list, launches, total = [], 0, 0.0
    while len(ms_list) < 3 or (total < min_total_s * 1e3 and len(ms_list) < max_repeats):
        eng.rewind()
        eng.run(min(ips, 50), trace=trace, logp=logp)           # leave the first (refresh) windows outside the region
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launches
        barrier()
        ev0.record()
        for _ in range(K):
            eng.run(ips, trace=trace, logp=logp)
        ev1.record()
        barrier()
        launches = eng.launches - l0
        ms = reduce_max(ev0.elapsed_time(ev1))
        ms_list.append(ms)
        total += ms
    check = float(logp[:, -1].mean().item())     # touch the result
    return ms_list, launches, check


def roofline_of(wl, value_per_gpu, launches, ms, nl_chain_steps, kernel, traffic=None):
    peak, peak_src = measured_peaks()
    o = wl['opts']
    bs = b_step(wl['d'], o['multitry'], o['snooker'], o['DEpairs'], o['history_thin'])
    launch_ms = ms / max(launches, 1)
    per_launch = nl_chain_steps / max(launches, 1)
    achieved = bs * per_launch / (launch_ms * 1e-3) / 1e9
    return dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak, traffic=traffic,
                peak_source=peak_src, kernel=kernel, bytes_per_chain_step=bs, chain_steps_per_launch=per_launch,
                launch_ms=launch_ms)


def bench_other_config(name, K, W):
    """BASELINE configs 3 / 4 / 5 on ONE GPU, device resident: value, roofline, dominant kernel, launch duration."""
    import torch
    from pydream_b200.engine import DreamEngine
    wl = WORKLOADS[name]
    N = wl['chains']
    tgt = make_target(wl['target'], wl['d'])
    hist, starts = synthetic_inputs(wl, N)
    ips, thin = wl['iters_per_step'], wl['opts']['history_thin']
    sync = lambda: torch.cuda.synchronize()
    out = dict(workload=wl['label'], ndim=wl['d'], nchains=N, archive_seed_rows=wl['nseed'], iters_per_step=ips,
               l2_policy='archive %.0f MB > 126 MB L2' % (hist.shape[0] * ((wl['d'] + 3) // 4 * 4) * 8 / 1e6))
    eng = DreamEngine(wl['d'], N, hist, starts, tgt, seed=SEED, record_decisions=False,
                      reserve_iters=(W + K + 1) * ips + 64, **wl['opts'])
    del hist
    ms_list, launches, check = timed_regions(eng, wl, K, W, sync, lambda x: x, min_total_s=0.3, max_repeats=7)
    ms = float(np.median(ms_list))
    value = N * K * ips / (ms * 1e-3)
    out.update(value=value, unit='chain-steps/s', ms_per_step=ms / K, steps=K, repeats=len(ms_list),
               region_ms=[round(x, 3) for x in ms_list], gpu_launches=launches,
               roofline=roofline_of(wl, value, launches, ms, N * K * ips, wl['kernel']), mean_final_logp=check)
    eng.close()
    del eng
    if name == 'c5':
        # the config as BASELINE words it: crossover adaptation during burn-in (one launch per iteration + the reduction
        # kernels) and Gelman-Rubin on the device trace
        hist, starts = synthetic_inputs(wl, N)
        T, B = 200, 40
        eng = DreamEngine(wl['d'], N, hist, starts, tgt, seed=SEED, record_decisions=True, adapt_crossover=True,
                          crossover_burnin=B, reserve_iters=T + 64, **wl['opts'])
        del hist
        trace = torch.empty((N, T, eng.ld), dtype=torch.float64, device=eng.device)
        logp = torch.empty((N, T), dtype=torch.float64, device=eng.device)
        dec = torch.empty((N, T), dtype=torch.int32, device=eng.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        sync()
        l0 = eng.launches
        ev[0].record()
        eng._advance(B + 1, trace, logp, dec)             # burn-in: one launch per iteration + the reduction kernels
        ev[1].record()
        l1 = eng.launches
        _advance_into(eng, trace, logp, dec, B + 1, T - B - 1)
        ev[2].record()
        rhat = eng.gelman_rubin(trace)
        ev[3].record()
        sync()
        burn_ms, rhat_ms = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
        out['burnin'] = dict(iterations=B + 1, chain_steps_per_s=N * (B + 1) / (burn_ms * 1e-3), ms_per_iteration=burn_ms / (B + 1),
                             gpu_launches=l1 - l0, adapted_cr_probs=[float(x) for x in eng.cr_probs.cpu().numpy()])
        out['rhat'] = dict(ms=rhat_ms, trace_iterations=T, gbs=8.0 * N * wl['d'] * (T - T // 2) / (rhat_ms * 1e-3) / 1e9,
                           max=float(rhat.max().item()))
        eng.close()
    return out


def _advance_into(eng, trace, logp, dec, t0, n):
    """n more iterations into rows [t0, t0+n) of full-length trace buffers (the C ABI takes an offset)."""
    import ctypes as C
    from pydream_b200 import _cabi
    tr = _cabi.Trace(trace=trace.data_ptr(), trace_logp=logp.data_ptr(), decisions=dec.data_ptr(), trace_iters=trace.shape[1],
                     trace_offset=t0)
    nl, rows = C.c_int64(0), C.c_int64(0)
    rc = eng.lib.dreamzs_run(C.byref(eng.cfg), C.byref(eng.st), C.byref(tr), eng.iter, n, eng.archive_rows, eng.count // eng.N,
                             None, _cabi.APPEND_HOOK(), None, None, eng._stream(), C.byref(nl), C.byref(rows))
    _cabi.check(rc, 'dreamzs_run')
    eng.launches += int(nl.value)
    eng.count = int(rows.value) - eng.nseed
    eng.iter += n


def sharded_check(group, world, rank, device):
    """A short sharded run against the identical single-GPU run on rank 0 (bit for bit), outside any timed region."""
    import torch
    import torch.distributed as dist
    from pydream_b200.engine import DreamEngine
    wl = WORKLOADS['c2']
    d, N, T = wl['d'], 56 * world, 45
    tgt = make_target('gaussian', d)
    rng = np.random.default_rng(99)
    hist = rng.uniform(-5, 15, size=(2 * N + 9, d))
    kw = dict(seed=4, snooker=.1, history_thin=10)
    eng = DreamEngine(d, N, hist, hist[:N].copy(), tgt, group=group, **kw)
    T1 = 17
    parts = [eng.run(T1), eng.run(T - T1)]          # the archive grows between the calls
    trace, logp, dec = (torch.cat([a[i] for a in parts], dim=1).contiguous() for i in range(3))
    torch.cuda.synchronize()
    eng.check_peers()
    transport = 'nvlink peer stores' if eng.peers is not None else 'nccl all-gather'
    Zloc = eng.Z[:eng.archive_rows].clone()
    gt = [torch.empty_like(trace) for _ in range(world)]
    gl = [torch.empty_like(logp) for _ in range(world)]
    gd = [torch.empty_like(dec) for _ in range(world)]
    gz = [torch.empty_like(Zloc) for _ in range(world)]
    dist.all_gather(gt, trace, group=group)
    dist.all_gather(gl, logp, group=group)
    dist.all_gather(gd, dec, group=group)
    dist.all_gather(gz, Zloc, group=group)
    eng.close()
    ok = None
    if rank == 0:
        one = DreamEngine(d, N, hist, hist[:N].copy(), tgt, **kw)
        t1, l1, d1 = one.run(T)
        torch.cuda.synchronize()
        Z1 = one.Z[:one.archive_rows]
        ok = bool(torch.equal(torch.cat(gt), t1) and torch.equal(torch.cat(gl), l1) and torch.equal(torch.cat(gd), d1)
                  and all(torch.equal(z, Z1) for z in gz))
    return ok, transport, dict(ndim=d, nchains=N, iterations=T)


def bench_gpu(args):
    import torch
    import torch.distributed as dist
    from pydream_b200.engine import DreamEngine
    from pydream_b200.core import run_dream
    from pydream_b200.parameters import FlatParam

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        group = dist.group.WORLD
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d' % args.gpus
    wl = WORKLOADS[args.workload]
    D = wl['d']
    N = total_chains(wl, world)
    tgt = make_target(wl['target'], D)
    hist, starts = synthetic_inputs(wl, N)
    K, W = args.steps, args.warmup
    ips, thin = wl['iters_per_step'], wl['opts']['history_thin']
    dev = torch.device('cuda', local)
    bg = start_reference_python_default() if (rank == 0 and world == 1 and not args.no_cpu) else None

    def barrier():
        if world > 1:
            dist.barrier(group)
        torch.cuda.synchronize()

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    # ---------------- device-resident timing (inputs already in HBM)
    eng = DreamEngine(D, N, hist, starts, tgt, seed=SEED, group=group, record_decisions=False,
                      reserve_iters=(W + K + 1) * ips + 64, **wl['opts'])
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
        time.sleep(0.12)
    ms_list, launches, acc_check = timed_regions(eng, wl, K, W, barrier, reduce_max)
    clocks = sampler.stop() if sampler else None
    ms = float(np.median(ms_list))
    value = N * K * ips / (ms * 1e-3)
    eng.check_peers()
    transport = 'nvlink peer stores' if eng.peers is not None else ('nccl all-gather' if world > 1 else 'none')
    Nl = eng.Nl
    eng.close()
    del eng

    # ---------------- end to end through the public API with host buffers
    Ke = args.e2e_iters
    pri = FlatParam(test_value=np.zeros(D))
    kw = dict(wl['opts'])
    kw['multitry'] = False if kw['multitry'] == 1 else kw['multitry']

    def e2e_run(hist_np, start_list):
        for _ in range(3):   # warm-up: same call, results dropped (the pinned result blocks return to torch's host cache)
            run_dream([pri], tgt, nchains=N, niterations=Ke, start=start_list, start_random=False, verbose=False,
                      history_file=hist_np, save_history=False, adapt_crossover=False, seed=SEED, group=group, **kw)
        barrier()
        t0 = time.perf_counter()
        run_dream([pri], tgt, nchains=N, niterations=Ke, start=start_list, start_random=False, verbose=False,
                  history_file=hist_np, save_history=False, adapt_crossover=False, seed=SEED, group=group, **kw)
        barrier()
        return reduce_max(time.perf_counter() - t0)

    # the caller's inputs live in pinned host memory (the archive seed is uploaded inside the timed region)
    hist_pinned = torch.empty(hist.shape, dtype=torch.float64, pin_memory=True)
    hist_pinned.numpy()[:] = hist
    start_list = [starts[c] for c in range(N)]
    e2e_s = e2e_run(hist_pinned.numpy(), start_list)
    e2e_value = N * Ke / e2e_s
    # SURVEY 8(d)'s own archive seed, max(10 d, 2 N) rows (the L2-busting 262144-row seed above costs a 210-MB upload)
    ns_small = max(10 * D, 2 * N)
    e2e_small_s = e2e_run(hist_pinned.numpy()[:ns_small], start_list)
    steps_e2e = Ke / ips
    h2d = (hist.nbytes + starts.nbytes) / steps_e2e
    d2h = Nl * ips * (D + 1) * 8

    check = dict(mean_final_logp=acc_check, logp_tol=LOGP_TOL)
    if world > 1:
        ok, tr2, shape = sharded_check(group, world, rank, dev)
        check.update(sharded_equals_single=ok, transport=tr2, sharded_check=shape)

    if rank == 0:
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp) and args.workload == 'c2':
            with open(tp) as f:
                tj = json.load(f)
                traffic = tj.get('dram_bytes_per_launch')
                if traffic is None and tj.get('dram_bytes_per_chain_step') is not None:   # per launch like `achieved`
                    traffic = tj['dram_bytes_per_chain_step'] * Nl * K * ips / max(launches, 1)
        line = dict(metric='chain-steps/sec', value=value, unit='chain-steps/s', n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms / K, higher_is_better=True, scaling=wl['scaling'], vs_baseline=None, dtype='f64',
                    data='synthetic', repeats=len(ms_list), region_ms=[round(x, 3) for x in ms_list],
                    timed_seconds=float(np.sum(ms_list)) * 1e-3,
                    config=config_of(wl, N, world,
                                     l2_policy='inputs larger than L2: archive >= %.0f MB, gathered rows are random'
                                               % (wl['nseed'] * ((D + 3) // 4 * 4) * 8 / 1e6),
                                     fused_iterations_per_launch=K * ips // max(launches, 1), e2e_iterations=Ke, archive_replication=transport,
                                     timing='median of `repeats` regions of K steps, archive rewound in between'),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit='chain-steps/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             seconds=e2e_s, iterations=Ke, steps=steps_e2e,
                             api='pydream_b200.core.run_dream (numpy in, numpy out)',
                             survey_seed=dict(archive_seed_rows=ns_small, value=N * Ke / e2e_small_s, seconds=e2e_small_s)),
                    gpu_launches=launches,
                    roofline=roofline_of(wl, value / world, launches, ms, Nl * K * ips, wl['kernel'], traffic),
                    check=check)
        if world == 1 and not args.no_other and args.workload == 'c2':
            others = {}
            for name in ('c3', 'c4', 'c5'):
                try:
                    others[name] = bench_other_config(name, max(2, min(K, 5)), max(1, min(W, 2)))
                except Exception as e:      # noqa: BLE001
                    others[name] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
                torch.cuda.empty_cache()
            line['other_configs'] = others
        if world == 1 and not args.no_cpu:
            cpu_value, cpu_dt, cpu_iters, nthreads = oracle_sample(wl, N, 12.0)
            cb = dict(value=cpu_value, unit='chain-steps/s', cores=nthreads, kind='port',
                      sample='%d iterations of %d chains (%.1f s), oracle/dreamzs_oracle.c on %d threads'
                             % (cpu_iters, N, cpu_dt, nthreads))
            rp = reference_python(host_threads())
            if bg is not None:
                try:
                    so, _ = bg.communicate(timeout=120)
                    rp.append(json.loads(so.strip().splitlines()[-1]))
                except Exception as e:      # noqa: BLE001
                    rp.append(dict(mode='default', unavailable='%s' % type(e).__name__))
            cb['reference_python'] = rp
            line['cpu_baseline'] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--e2e-iters', type=int, default=2000)
    ap.add_argument('--cpu-budget', type=float, default=60.0, help='seconds of CPU work of the reference arm')
    ap.add_argument('--no-other', action='store_true', help='skip other_configs')
    ap.add_argument('--no-cpu', action='store_true', help='skip cpu_baseline')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        bench_reference(args)
    else:
        bench_gpu(args)


if __name__ == '__main__':
    main()
