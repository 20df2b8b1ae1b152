"""world_size-2 gloo test of the host-side sharding logic: in-place all-gather of newly appended archive
rows into the replicated archive tail, and the all-reduced adaptation partials."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pydream_b200.collectives import allgather_rows, allreduce_sum
    N, ld, M = 6, 4, 5
    Nl = N // world
    Z = torch.zeros((M + N, ld), dtype=torch.float64)
    Z[:M] = torch.arange(M * ld, dtype=torch.float64).reshape(M, ld)
    c0 = rank * Nl
    for c in range(c0, c0 + Nl):          # what the kernel writes: row M + global chain id
        Z[M + c] = 100 + c
    allgather_rows(Z[M:M + N], c0, Nl, dist.group.WORLD)
    part = torch.tensor([1.0 + rank, 10.0 * (rank + 1)], dtype=torch.float64)
    allreduce_sum(part, dist.group.WORLD)
    if rank == 0:
        torch.save(dict(Z=Z, part=part), out)
    dist.barrier()
    dist.destroy_process_group()


def test_archive_allgather_layout(tmp_path):
    out = str(tmp_path / 'r0.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    Z = r['Z'].numpy()
    assert np.array_equal(Z[5:, 0], 100 + np.arange(6))      # chain order, identical to the 1-GPU layout
    assert np.array_equal(Z[:5].reshape(-1), np.arange(20))
    assert list(r['part'].numpy()) == [3.0, 30.0]
