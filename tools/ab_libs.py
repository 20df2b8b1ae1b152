"""Same-process A/B of builds of libdreamzs.so on the C2 workload (window kernel): interleaved repetitions, CUDA-event
timing, and a check that every build takes the decisions / produces the log-posteriors of the first one.
usage: python tools/ab_libs.py [--iters 2000] [--reps 3] name=path/to/lib.so [name=path ...]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--chains', type=int, default=1024)
    ap.add_argument('--dim', type=int, default=100)
    ap.add_argument('--nseed', type=int, default=262144)
    ap.add_argument('libs', nargs='+')
    a = ap.parse_args()
    import torch
    from pydream_b200 import _cabi, targets
    from pydream_b200.engine import DreamEngine
    libs = [x.split('=', 1) for x in a.libs]
    rng = np.random.default_rng(1)
    d, N = a.dim, a.chains
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(a.nseed, d))

    def engine(path, **kw):
        _cabi._lib, _cabi.LIB_PATH = None, os.path.abspath(path)
        return DreamEngine(d, N, hist, hist[:N], tgt, seed=0, snooker=.1, history_thin=10, **kw)

    ref = None
    for name, path in libs:      # parity between builds: 61 iterations with decisions
        eng = engine(path, record_decisions=True)
        tr, lp, dec = eng.run(61)
        torch.cuda.synchronize()
        got = (dec.cpu().numpy(), lp.cpu().numpy())
        if ref is None:
            ref = got
        else:
            same_dec = np.array_equal(got[0], ref[0])
            err = np.abs(got[1] - ref[1]).max()
            print('parity %s vs %s: decisions %s, max |dlogp| %.3e' % (name, libs[0][0], 'identical' if same_dec else 'DIFFER', err))
    res = {name: [] for name, _ in libs}
    for rep in range(a.reps):
        for name, path in libs:
            eng = engine(path, record_decisions=False)
            eng.run(111)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(a.iters)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(1e3 * e0.elapsed_time(e1) / a.iters)
            del eng
    for name, _ in libs:
        v = res[name]
        print('%-12s us/iteration: %s  best %.3f  -> %.1f M chain-steps/s' % (name, ' '.join('%.3f' % x for x in v), min(v), N / min(v)))


if __name__ == '__main__':
    main()
