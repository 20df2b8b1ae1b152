"""CPU baseline of SURVEY.md 8(d): the UNMODIFIED reference's own `pydream.core.run_dream` (one OS process per chain,
pydream/core.py:66-86, 307-310) timed on this host.  Imports pydream from /root/reference when it exists (build
container), else from baseline/_ref (the unchanged copy installed with pip --target, which travels to the GPU box).

    python tools/ref_python_bench.py --nchains 8 --niter 1000 [--mode nosleep|default] [--dim 100] [--multitry 1]

mode nosleep: crossover_burnin > niterations, so the burn-in "barrier" of Dream.py:385-415 (time.sleep(30) polling
              loop + time.sleep(10)) stays outside the run: pure compute.  adapt_crossover as given (default False).
mode default: the reference's defaults (crossover_burnin = niter/10, adapt_crossover=True): includes >= 40 s of sleeps.
The target is the reference's own example likelihood (dream_ex_ndim_gaussian.py:30-52) at dimension --dim, FlatParam
prior, archive seeded from a history file, chain c starts on seed row c (:54, :65).  OMP_NUM_THREADS=1 (set before
numpy is imported).  Prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

os.environ.setdefault('OMP_NUM_THREADS', '1')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
os.environ.setdefault('MKL_NUM_THREADS', '1')

import numpy as np   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _root in ('/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
    if os.path.isdir(os.path.join(_root, 'pydream')):
        sys.path.insert(0, _root)
        REF = _root
        break
else:
    REF = None

invC = None
log_F = 0.0


def likelihood(param_vec):
    # dream_ex_ndim_gaussian.py:49-52
    logp = log_F - .5 * np.sum(param_vec*np.dot(invC, param_vec))
    return logp


def main():
    global invC, log_F
    ap = argparse.ArgumentParser()
    ap.add_argument('--nchains', type=int, default=3)
    ap.add_argument('--niter', type=int, default=1000)
    ap.add_argument('--dim', type=int, default=100)
    ap.add_argument('--mode', default='nosleep', choices=['nosleep', 'default'])
    ap.add_argument('--multitry', type=int, default=1)
    ap.add_argument('--adapt', type=int, default=0)
    a = ap.parse_args()
    if REF is None:
        print(json.dumps(dict(unavailable='neither /root/reference nor baseline/_ref holds pydream')))
        return
    from pydream.core import run_dream
    from pydream.parameters import FlatParam
    d = a.dim
    A = .5 * np.identity(d) + .5 * np.ones((d, d))
    idx = np.arange(1, d + 1, dtype=np.float64)
    C = A * np.sqrt(np.outer(idx, idx))
    invC = np.linalg.inv(C)
    log_F = 0 if d > 150 else np.log(((2 * np.pi)**(-d/2))*np.linalg.det(C)**(- 1./2))
    rng = np.random.default_rng(1234)
    nseed = max(10 * d, 2 * a.nchains)
    m = rng.uniform(-5, 15, size=(nseed, d))
    starts = [m[c] for c in range(a.nchains)]
    params = FlatParam(test_value=np.zeros(d))
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix='ref_python_bench_')
    os.chdir(tmp)
    try:
        np.save('seed.npy', m)
        kw = dict(niterations=a.niter, nchains=a.nchains, start=starts, start_random=False, save_history=False,
                  history_file='seed.npy', verbose=False, multitry=(a.multitry if a.multitry > 1 else False))
        if a.mode == 'nosleep':
            kw.update(crossover_burnin=10 ** 9, adapt_crossover=bool(a.adapt))
        t0 = time.perf_counter()
        sampled, log_ps = run_dream([params], likelihood, **kw)
        dt = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    assert len(sampled) == a.nchains and sampled[0].shape == (a.niter, d)
    print(json.dumps(dict(impl='pydream.core.run_dream (unmodified, %s)' % REF, mode=a.mode, nchains=a.nchains, niter=a.niter,
                          ndim=d, multitry=a.multitry, seconds=dt, chain_steps_per_s=a.nchains * a.niter / dt,
                          host_cores=os.cpu_count(), omp_num_threads=os.environ.get('OMP_NUM_THREADS'),
                          mean_final_logp=float(np.mean([lp[-1, 0] for lp in log_ps])))))


if __name__ == '__main__':
    main()
