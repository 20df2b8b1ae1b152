"""A/B of builds of libdreamzs.so on the C2 workload (window kernel): one child process per (build, repetition),
interleaved (a second library of the same kernels cannot be loaded into one process: measured, the launch fails),
CUDA-event timing, and a check that every build takes the decisions / produces the log-posteriors of the first.
usage: python tools/ab_libs.py [--iters 2000] [--reps 2] name=path/to/lib.so [name=path ...]
Round-1 use: base vs -DDZ_GW_GROUPS=4."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(a):
    import numpy as np
    import torch
    from pydream_b200 import targets
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(1)
    d, N = a.dim, a.chains
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(a.nseed, d))
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=0, snooker=.1, history_thin=10, record_decisions=True)
    tr, lp, dec = eng.run(61)
    torch.cuda.synchronize()
    h = hashlib.sha1(dec.cpu().numpy().tobytes()).hexdigest()[:12]
    hl = hashlib.sha1(lp.cpu().numpy().tobytes()).hexdigest()[:12]
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=0, snooker=.1, history_thin=10, record_decisions=False)
    eng.run(111)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(a.iters)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps(dict(us_per_iter=1e3 * e0.elapsed_time(e1) / a.iters, decisions=h, logp=hl)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--chains', type=int, default=1024)
    ap.add_argument('--dim', type=int, default=100)
    ap.add_argument('--nseed', type=int, default=262144)
    ap.add_argument('--child', action='store_true')
    ap.add_argument('libs', nargs='*')
    a = ap.parse_args()
    if a.child:
        return child(a)
    libs = [x.split('=', 1) for x in a.libs]
    res = {name: [] for name, _ in libs}
    for rep in range(a.reps):
        for name, path in libs:
            env = dict(os.environ, DREAMZS_LIB=os.path.abspath(path))
            out = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', '--iters', str(a.iters), '--chains', str(a.chains),
                                  '--dim', str(a.dim), '--nseed', str(a.nseed)], env=env, capture_output=True, text=True, timeout=300)
            if out.returncode != 0:
                print('%s failed:\n%s' % (name, out.stderr[-2000:]))
                continue
            res[name].append(json.loads(out.stdout.strip().splitlines()[-1]))
    first = res[libs[0][0]][0] if res[libs[0][0]] else None
    for name, _ in libs:
        v = [r['us_per_iter'] for r in res[name]]
        if not v:
            continue
        same = first is not None and all(r['decisions'] == first['decisions'] for r in res[name])
        bits = first is not None and all(r['logp'] == first['logp'] for r in res[name])
        print('%-12s us/iteration: %s  best %.3f -> %.1f M chain-steps/s; decisions %s, logp %s'
              % (name, ' '.join('%.3f' % x for x in v), min(v), a.chains / min(v), 'identical' if same else 'DIFFER',
                 'bit-identical' if bits else 'differs in rounding'))


if __name__ == '__main__':
    main()
