"""Minimal driver for ncu: runs a few fused-step launches of the C2 workload (no CPU arm, no e2e, no
subprocesses).  usage: python tools/profile_step.py [--generic] [--iters 40] [--chains 1024] [--dim 100]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--generic', action='store_true')
    ap.add_argument('--old', action='store_true', help='the round-1 window kernel (carried y = invC x) instead of the whitened one')
    ap.add_argument('--thin', type=int, default=10)
    ap.add_argument('--nopersist', action='store_true', help='one launch per window instead of one persistent launch per run')
    ap.add_argument('--iters', type=int, default=41)
    ap.add_argument('--chains', type=int, default=1024)
    ap.add_argument('--dim', type=int, default=100)
    ap.add_argument('--nseed', type=int, default=262144)
    ap.add_argument('--snooker', type=float, default=.1)
    ap.add_argument('--target', default='gaussian')
    ap.add_argument('--multitry', type=int, default=1)
    ap.add_argument('--time', action='store_true', help='print CUDA-event timing of the run')
    ap.add_argument('--phases', action='store_true', help='print the clock64 phase stamps of CTA 0 of the last window-kernel launch')
    a = ap.parse_args()
    import torch
    from pydream_b200 import targets
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(1)
    d, N = a.dim, a.chains
    if a.target == 'gaussian':
        tgt = targets.CorrelatedGaussian.benchmark(d)
        hist = rng.uniform(-5, 15, size=(a.nseed, d))
    elif a.target == 'mixture':
        tgt = targets.BimodalMixture.benchmark(d)
        hist = rng.normal(size=(a.nseed, d))
    else:
        tgt = targets.Banana(d)
        hist = rng.normal(size=(a.nseed, d))
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=0, snooker=a.snooker, history_thin=a.thin, multitry=a.multitry,
                      record_decisions=False, generic_kernel=a.generic, whitened=not a.old, persistent=not a.nopersist)
    eng.run(11)
    torch.cuda.synchronize()
    if a.phases:
        import ctypes
        buf = torch.zeros(96, dtype=torch.int64, device='cuda')
        eng.lib.dreamzs_debug_set_phase_buffer(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launches
    e0.record()
    eng.run(a.iters)
    e1.record()
    torch.cuda.synchronize()
    if a.time:
        ms = e0.elapsed_time(e1)
        print('%s d=%d N=%d: %d iterations in %.3f ms -> %.2f us/iter, %.1f M chain-steps/s, %d launches'
              % ('generic' if a.generic else 'old window kernel' if a.old else 'auto', d, N, a.iters, ms, 1e3 * ms / a.iters, N * a.iters / ms / 1e3,
                 eng.launches - l0))


    if a.phases:
        t = buf.cpu().numpy()
        acc = t[60:70].copy()
        t = t[:60]
        if acc[0]:
            nb = float(acc[0])
            print('all CTAs, cycles per batch: columns %.0f, products %.0f, chains (warp 0) %.0f, wait for other warps %.0f, loop top %.0f; '
                  'row waits: %d, %.0f cycles each' % (acc[1] / nb, acc[2] / nb, acc[3] / nb, acc[4] / nb, acc[5] / nb, acc[7], acc[6] / max(acc[7], 1)))
        t0 = int(t[t != 0].min()) if (t != 0).any() else 0
        for name, lo in (('C warp 0 (iteration ends)', 0), ('V warps (fill start, fill end, next pre end)', 20), ('M', 40)):
            seg = [int(v - t0) for v in t[lo:lo + 20] if v != 0]
            if seg:
                print('%s: %s' % (name, seg))
                print('   deltas: %s' % [seg[i + 1] - seg[i] for i in range(len(seg) - 1)])


if __name__ == '__main__':
    main()
