#!/usr/bin/env python
"""Benchmark of the MT-DREAM(ZS) step path on B200 (BASELINE.json metric: chain-steps/sec).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host threads

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): 100-D correlated Gaussian
(dream_ex_ndim_gaussian.py covariance), 1024 chains per GPU, FlatParam prior, reference default options
(snooker .1, DEpairs 1, nCR 3, history_thin 10, multitry off), crossover adaptation off so that the timed
region is the steady-state step.  One "step" = one sampler iteration of every chain (1024 chain-steps per
GPU); a launch of the dense-Gaussian window kernel fuses the 10 iterations of one history_thin window.
The archive is pre-seeded with 262144 rows (210 MB > the 126 MB L2) so every timed gather works on an input
larger than L2.  Scaling over GPUs is weak: 1024 chains per GPU, archive replicated, the rows appended every
history_thin iterations reach the other replicas as NVLink peer stores from inside the step kernel (NCCL
all-gather when peer mappings are unavailable).

`value`  : K steps through DreamEngine.run with everything resident in HBM, CUDA events, max over ranks.
`e2e`    : the same metric through pydream_b200.core.run_dream with host (numpy) inputs and outputs: archive
           seed + starts uploaded, every sample and log-posterior copied back, inside the timed region.
`roofline`: algorithmic HBM bytes per chain-step x chain-steps per launch / launch duration,
           against the measured copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline` (N=1): oracle/dreamzs_oracle.c (a C port of the reference's step) on the host threads, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, CHAINS_PER_GPU, NSEED = 100, 1024, 262144
OPTS = dict(snooker=.1, history_thin=10, DEpairs=1, nCR=3, multitry=1, p_gamma_unity=.2, lamb=.05, zeta=1e-12)
SEED = 0


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


def synthetic_inputs(nchains, rng_seed=1234):
    """Archive seed ~ U(-5,15)^d (the example's Latin-hypercube box, dream_ex_ndim_gaussian.py:17-26,45);
    chain c starts at seed row c (:54)."""
    rng = np.random.default_rng(rng_seed)
    hist = rng.uniform(-5, 15, size=(NSEED, D))
    return hist, hist[:nchains].copy()


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


def run_oracle_cpu(nchains, max_iter, nthreads, warmup=3, budget_s=15.0):
    """The CPU arm: oracle/dreamzs_oracle.c (a port of the reference's step path) on host threads.
    Runs at most `max_iter` iterations in growing chunks and stops once `budget_s` seconds have been
    spent (bounded sample).  Returns (chain-steps/s, seconds, iterations timed)."""
    from oracle import c_oracle
    from pydream_b200 import targets
    tgt = targets.CorrelatedGaussian.benchmark(D)
    hist, starts = synthetic_inputs(nchains)
    s = c_oracle.OracleSampler(D, nchains, hist, starts, tgt.kind, tgt.table(), seed=SEED, nthreads=nthreads,
                               capacity_rows=NSEED + ((warmup + max_iter) // OPTS['history_thin'] + 2) * nchains, **OPTS)
    if warmup:
        s.run(warmup)
    done, dt, chunk = 0, 0.0, 5
    while done < max_iter and dt < budget_s:
        n = min(chunk, max_iter - done)
        t0 = time.perf_counter()
        s.run(n)
        dt += time.perf_counter() - t0
        done += n
        rate = done / dt
        chunk = int(max(5, min(max_iter - done, rate * max(budget_s - dt, 0.0) * 0.5 + 1)))
    return nchains * done / dt, dt, done


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


def best_cpu_arm(nchains, max_iter, budget_s):
    """Try the full thread count and a few smaller ones on a short sample (oversubscribed or throttled hosts
    run slower with more threads), then spend the budget on the best."""
    nmax = host_threads()
    cands = sorted({nmax, max(1, nmax // 2), min(nmax, 32), min(nmax, 16), min(nmax, 8)}, reverse=True)
    best, best_rate = cands[-1], 0.0
    for n in cands:
        rate, _, _ = run_oracle_cpu(nchains, 40, n, warmup=2, budget_s=2.0)
        if rate > best_rate:
            best, best_rate = n, rate
    rate, dt, done = run_oracle_cpu(nchains, max_iter, best, warmup=3, budget_s=budget_s)
    return rate, dt, done, best


def bench_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    nchains = CHAINS_PER_GPU
    # bounded sample: at most the requested steps, at most about 40 s of CPU work
    value, dt, steps, nthreads = best_cpu_arm(nchains, args.steps, 40.0)
    warm = 3
    sample = '%d of the requested %d iterations of %d chains, %d threads' % (steps, args.steps, nchains, nthreads)
    line = dict(impl='reference', metric='chain-steps/sec', value=value, unit='chain-steps/s', n_gpus=args.gpus,
                steps=steps, warmup=warm, ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload='C2: 100-D correlated Gaussian, 1024 chains, MT-DREAM(ZS) step (CPU arm: C port of '
                                     'pydream Dream.astep, lock-step sweep split over host threads)',
                            ndim=D, nchains=nchains, archive_seed_rows=NSEED, **OPTS),
                cpu_baseline=dict(value=value, unit='chain-steps/s', cores=nthreads, kind='port', sample=sample),
                e2e=dict(value=value, unit='chain-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def bench_gpu(args):
    import torch
    import torch.distributed as dist
    from pydream_b200 import targets
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
    N = CHAINS_PER_GPU * world
    tgt = targets.CorrelatedGaussian.benchmark(D)
    hist, starts = synthetic_inputs(N)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier(group)
        torch.cuda.synchronize()

    # ---------------- device-resident timing (inputs already in HBM)
    eng = DreamEngine(D, N, hist, starts, tgt, seed=SEED, group=group, record_decisions=False, **OPTS)
    eng._ensure_capacity(NSEED + ((W + K) // OPTS['history_thin'] + 2) * N)
    wtrace = torch.empty((eng.Nl, W, eng.ld), dtype=torch.float64, device=eng.device)
    wlogp = torch.empty((eng.Nl, W), dtype=torch.float64, device=eng.device)
    eng.run(W, trace=wtrace, logp=wlogp)
    del wtrace, wlogp
    trace = torch.empty((eng.Nl, K, eng.ld), dtype=torch.float64, device=eng.device)
    logp = torch.empty((eng.Nl, K), dtype=torch.float64, device=eng.device)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
        time.sleep(0.12)
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    eng.run(K, trace=trace, logp=logp)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launches - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms = float(t.item())
    value = N * K / (ms * 1e-3)
    acc_check = float(logp[:, -1].mean().item())   # touch the result
    eng.check_peers()
    transport = 'nvlink peer stores' if eng.peers is not None else ('nccl all-gather' if world > 1 else 'none')
    eng.close()
    del trace, logp

    # ---------------- end to end through the public API with host buffers
    Ke = min(K, args.e2e_steps)
    pri = FlatParam(test_value=np.zeros(D))
    start_list = [starts[c] for c in range(N)]
    kw = dict(OPTS)
    kw['multitry'] = False
    # the caller's inputs live in pinned host memory (the archive seed is uploaded inside the timed region)
    hist_pinned = torch.empty(hist.shape, dtype=torch.float64, pin_memory=True)
    hist_pinned.numpy()[:] = hist
    hist = hist_pinned.numpy()
    for _ in range(2):   # warm-up: same call, results dropped (the pinned result blocks return to torch's host cache)
        run_dream([pri], tgt, nchains=N, niterations=Ke, start=start_list, start_random=False, verbose=False,
                  history_file=hist, save_history=False, adapt_crossover=False, seed=SEED, group=group, **kw)
    barrier()
    t0 = time.perf_counter()
    sp, lps = run_dream([pri], tgt, nchains=N, niterations=Ke, start=start_list, start_random=False, verbose=False,
                        history_file=hist, save_history=False, adapt_crossover=False, seed=SEED, group=group, **kw)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    e2e_s = float(t.item())
    e2e_value = N * Ke / e2e_s
    h2d = (hist.nbytes + starts.nbytes) / Ke
    d2h = eng.Nl * (D + 1) * 8

    if rank == 0:
        peak, peak_src = measured_peaks()
        bs = b_step(D, OPTS['multitry'], OPTS['snooker'], OPTS['DEpairs'], OPTS['history_thin'])
        thin = OPTS['history_thin']
        per_launch_steps = eng.Nl * K / max(launches, 1)
        launch_ms = ms / max(launches, 1)
        achieved = bs * per_launch_steps / (launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get('dram_bytes_per_launch')
        cpu_value = None
        if world == 1:
            cpu_value, cpu_dt, cpu_iters, nthreads = best_cpu_arm(CHAINS_PER_GPU, args.cpu_steps, 12.0)
        line = dict(metric='chain-steps/sec', value=value, unit='chain-steps/s', n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms / K, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                    data='synthetic',
                    config=dict(workload='C2: 100-D correlated Gaussian (dense precision), 1024 chains per GPU, fused '
                                         'DE/snooker proposal + logp + accept kernel',
                                ndim=D, nchains=N, chains_per_gpu=CHAINS_PER_GPU, archive_seed_rows=NSEED,
                                l2_policy='inputs larger than L2: archive >= 210 MB, gathered rows are random',
                                fused_iterations_per_launch=thin, e2e_steps=Ke, archive_replication=transport, **OPTS),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit='chain-steps/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             seconds=e2e_s, steps=Ke, api='pydream_b200.core.run_dream (numpy in, numpy out)'),
                    gpu_launches=launches,
                    roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                                  traffic=traffic, peak_source=peak_src, kernel='dreamzs_gwin_kernel<7>',
                                  bytes_per_chain_step=bs, chain_steps_per_launch=per_launch_steps,
                                  launch_ms=launch_ms),
                    check=dict(mean_final_logp=acc_check))
        if cpu_value is not None:
            line['cpu_baseline'] = dict(value=cpu_value, unit='chain-steps/s', cores=nthreads, kind='port',
                                        sample='%d iterations of %d chains (%.1f s), oracle/dreamzs_oracle.c on %d threads'
                                               % (cpu_iters, CHAINS_PER_GPU, cpu_dt, nthreads))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20000)
    ap.add_argument('--warmup', type=int, default=1000)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--e2e-steps', type=int, default=2000)
    ap.add_argument('--cpu-steps', type=int, default=2000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        bench_reference(args)
    else:
        bench_gpu(args)


if __name__ == '__main__':
    main()
