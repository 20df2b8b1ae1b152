"""C2 workload on N GPUs (one process per GPU, torchrun), device-timed, with the window kernel's phase counters of every rank.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/profile_multi.py [--iters 20010] [--solo]
--solo: every rank runs its own unsharded 1024-chain sampler (no peers): what the box gives N processes side by side."""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20010)
    ap.add_argument('--chains', type=int, default=1024, help='per GPU')
    ap.add_argument('--dim', type=int, default=100)
    ap.add_argument('--nseed', type=int, default=262144)
    ap.add_argument('--solo', action='store_true')
    ap.add_argument('--phases', action='store_true')
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    from pydream_b200 import targets
    from pydream_b200.engine import DreamEngine
    rng = np.random.default_rng(1)
    d = a.dim
    N = a.chains if a.solo else a.chains * world
    tgt = targets.CorrelatedGaussian.benchmark(d)
    hist = rng.uniform(-5, 15, size=(a.nseed, d))
    eng = DreamEngine(d, N, hist, hist[:N], tgt, seed=0, snooker=.1, history_thin=10, record_decisions=False,
                      group=None if a.solo else dist.group.WORLD, reserve_iters=a.iters + 11)
    eng.run(11)
    torch.cuda.synchronize()
    buf = None
    if a.phases:
        buf = torch.zeros(96, dtype=torch.int64, device='cuda')
        eng.lib.dreamzs_debug_set_phase_buffer(ctypes.c_void_p(buf.data_ptr()))
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(a.iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    msg = 'rank %d: %d iterations in %.3f ms -> %.2f us/iter, %.1f M chain-steps/s per GPU' % (rank, a.iters, ms, 1e3 * ms / a.iters, a.chains * a.iters / ms / 1e3)
    if buf is not None:
        acc = buf.cpu().numpy()[60:70]
        nb = max(float(acc[0]), 1.0)
        msg += '\n   cycles per batch: columns %.0f, products %.0f, chains (warp 0) %.0f, wait for other warps %.0f, loop top %.0f; row waits: %d, %.0f cycles each' % (
            acc[1] / nb, acc[2] / nb, acc[3] / nb, acc[4] / nb, acc[5] / nb, acc[7], acc[6] / max(acc[7], 1))
    for r in range(world):
        if r == rank:
            print(msg, flush=True)
        dist.barrier()
    if not a.solo:
        eng.check_peers()
        eng.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
