"""Timing of the auxiliary kernels of the path: Gelman-Rubin (convergence.py:3-20) and one burn-in adaptation sweep
(Dream.py:451-499) at the C5 shape, reported as achieved HBM GB/s of their algorithmic bytes."""
import os, sys, time
ONLY_ADAPT = '--adapt-only' in sys.argv
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydream_b200 import targets
from pydream_b200.engine import DreamEngine, gelman_rubin_device

dev = torch.device('cuda:0')
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# Gelman-Rubin: N chains x T iterations x d (C5: 65536 x 1000 x 50 -> ld 52)
for (N, T, d) in (() if ONLY_ADAPT else ((65536, 1000, 50), (1024, 10000, 100))):
    ld = (d + 3) // 4 * 4
    trace = torch.randn((N, T, ld), dtype=torch.float64, device=dev)
    ms = timed(lambda: gelman_rubin_device(trace, d))
    alg = 8.0 * N * d * (T - T // 2)
    print('Gelman_Rubin N=%d T=%d d=%d: %.3f ms, algorithmic %.2f GB -> %.0f GB/s' % (N, T, d, ms, alg / 1e9, alg / ms / 1e6))
    del trace

# adaptation sweep at C5 shape on 1 GPU: 8192 chains, d=50 (burn-in iteration = step launch + 4 reduction kernels)
rng = np.random.default_rng(0)
N, d = 8192, 50
tgt = targets.CorrelatedGaussian.benchmark(d)
hist = rng.uniform(-5, 15, size=(2 * N + 100, d))
eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=0, snooker=.1, history_thin=10, adapt_crossover=True, crossover_burnin=10 ** 6)
eng.run(30); torch.cuda.synchronize()
t0 = time.perf_counter(); l0 = eng.launches
eng.run(200); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('burn-in with CR adaptation, N=%d d=%d: %.1f us per iteration (%d launches per iteration), %.1f M chain-steps/s'
      % (N, d, 1e6 * dt / 200, (eng.launches - l0) // 200, N * 200 / dt / 1e6))
