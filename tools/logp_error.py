"""Error of the window kernel's incremental log-posterior against the direct evaluation of the generic kernel
(same draws, same decisions): max absolute and max relative difference over a run of the C2 shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydream_b200 import targets
from pydream_b200.engine import DreamEngine

d, N, T = 100, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(5)
tgt = targets.CorrelatedGaussian.benchmark(d)
hist = rng.uniform(-5, 15, size=(8192, d))
kw = dict(seed=3, snooker=.1, history_thin=10)
a = DreamEngine(d, N, hist, hist[:N], tgt, **kw).run(T)
b = DreamEngine(d, N, hist, hist[:N], tgt, generic_kernel=True, **kw).run(T)
la, lb = a[1].cpu().numpy(), b[1].cpu().numpy()
same = bool((a[2] == b[2]).all().item())
err = np.abs(la - lb)
print('decisions identical: %s; |logp| median %.1f; max abs err %.3e (at |logp| %.1f); max rel err %.3e; abs err in the last 500 iterations: max %.3e median %.3e'
      % (same, np.median(np.abs(lb)), err.max(), np.abs(lb).flat[err.argmax()], (err / np.maximum(1, np.abs(lb))).max(),
         err[:, -500:].max(), np.median(err[:, -500:])))
