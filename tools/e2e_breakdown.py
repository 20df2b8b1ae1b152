"""Where the end-to-end time of run_dream goes (C2 workload): PCIe bandwidths, then a timed call with stage stamps."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pydream_b200 import targets
from pydream_b200.core import run_dream
from pydream_b200.parameters import FlatParam

D, N, NSEED, K = 100, 1024, 262144, 2000
dev = torch.device('cuda:0')
# raw PCIe
x = torch.empty(1 << 28, dtype=torch.uint8, device=dev)   # 256 MB
h = torch.empty(1 << 28, dtype=torch.uint8, pin_memory=True)
for name, fn in (('D2H pinned 256MB', lambda: h.copy_(x, non_blocking=True)), ('H2D pinned 256MB', lambda: x.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 4
    print('%s: %.1f GB/s' % (name, (1 << 28) / dt / 1e9))
# 2-D pitched D2H like the trace chunks
from pydream_b200 import _cabi
import ctypes as C
lib = _cabi.load()
src = torch.empty((N, 256, D), dtype=torch.float64, device=dev)
dst = torch.empty((N, K, D), dtype=torch.float64, pin_memory=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for t in range(0, K - 255, 256):
    lib.dreamzs_copy_d2h_2d(C.c_void_p(dst.data_ptr() + t * D * 8), K * D * 8, C.c_void_p(src.data_ptr()), 256 * D * 8, 256 * D * 8, N, s)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print('pitched D2H (1024 rows x 204800 B per chunk): %.1f GB/s' % ((K // 256) * 256 * N * D * 8 / dt / 1e9))

rng = np.random.default_rng(0)
hist_p = torch.empty((NSEED, D), dtype=torch.float64, pin_memory=True)
hist = hist_p.numpy(); hist[:] = rng.uniform(-5, 15, size=(NSEED, D))
starts = [hist[c].copy() for c in range(N)]
tgt = targets.CorrelatedGaussian.benchmark(D)
kw = dict(snooker=.1, history_thin=10, DEpairs=1, nCR=3, multitry=False, p_gamma_unity=.2, lamb=.05, zeta=1e-12)
os.environ['DREAMZS_TIMING'] = '1'
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sp, lp = run_dream([FlatParam(test_value=np.zeros(D))], tgt, nchains=N, niterations=K, start=starts, start_random=False, verbose=False,
                       history_file=hist, save_history=False, adapt_crossover=False, seed=0, **kw)
    dt = time.perf_counter() - t0
    print('run_dream %d: %.1f ms -> %.1f M chain-steps/s' % (rep, 1e3 * dt, N * K / dt / 1e6))
    del sp, lp
os.environ.pop('DREAMZS_TIMING')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
sp, lp = run_dream([FlatParam(test_value=np.zeros(D))], tgt, nchains=N, niterations=K, start=starts, start_random=False, verbose=False,
                   history_file=hist, save_history=False, adapt_crossover=False, seed=0, **kw)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
