import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from pydream_b200 import targets
from pydream_b200.core import run_dream
from pydream_b200.parameters import FlatParam
D, NSEED, K = 100, 262144, 2000
N = 1024 * dist.get_world_size()
rng = np.random.default_rng(0)
hp = torch.empty((NSEED, D), dtype=torch.float64, pin_memory=True); hist = hp.numpy(); hist[:] = rng.uniform(-5, 15, size=(NSEED, D))
starts = [hist[c].copy() for c in range(N)]
tgt = targets.CorrelatedGaussian.benchmark(D)
kw = dict(snooker=.1, history_thin=10, DEpairs=1, nCR=3, multitry=False, p_gamma_unity=.2, lamb=.05, zeta=1e-12)
if dist.get_rank() == 0: os.environ['DREAMZS_TIMING'] = '1'
for rep in range(5):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    sp, lp = run_dream([FlatParam(test_value=np.zeros(D))], tgt, nchains=N, niterations=K, start=starts, start_random=False, verbose=False,
                       history_file=hist, save_history=False, adapt_crossover=False, seed=0, group=dist.group.WORLD, **kw)
    dist.barrier(); dt = time.perf_counter() - t0
    if dist.get_rank() == 0: print('rep %d: %.1f ms -> %.1f M chain-steps/s' % (rep, 1e3 * dt, N * K / dt / 1e6), flush=True)
    del sp, lp
dist.destroy_process_group()
