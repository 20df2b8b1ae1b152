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


def _shard_worker(rank, world, port, out):
    """One rank of a sharded lock-step run on CPU: the C oracle steps this rank's block of chains (global chain ids in
    the random streams, appended rows at `rows so far + global chain id`), and after every appending iteration the
    product's own exchange step (pydream_b200.collectives.allgather_rows, the NCCL path of the engine) brings in the
    other rank's rows -- the schedule of DreamEngine's append hook, with gloo in place of NCCL."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import c_oracle
    from pydream_b200 import targets
    from pydream_b200.collectives import allgather_rows
    d, N, T, thin = 10, 8, 24, 3
    Nl, c0 = N // world, rank * (N // world)
    rng = np.random.default_rng(7)
    hist = rng.normal(size=(40, d))
    tgt = targets.BimodalMixture.benchmark(d)
    kw = dict(seed=21, snooker=.2, history_thin=thin, multitry=3)
    s = c_oracle.OracleSampler(d, Nl, hist, hist[c0:c0 + Nl], tgt.kind, tgt.table(), chain_begin=c0, nchains_global=N,
                               capacity_rows=40 + N * (T // thin), **kw)
    states, logps, decs = [], [], []
    for t in range(T):
        M = s.nseed + s.count.value
        o = s.run(1)
        if t % thin == 0:
            block = torch.from_numpy(s.Z)[M:M + N]          # this append's rows in every rank's replica
            allgather_rows(block, c0, Nl, dist.group.WORLD)
        states.append(o['states'][0]); logps.append(o['logp'][0]); decs.append(o['decisions'][0])
    torch.save(dict(states=np.stack(states), logp=np.stack(logps), dec=np.stack(decs), hist=s.history_flat), out % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_equals_single_process(tmp_path):
    """results must not depend on how the chains are sharded."""
    from oracle import c_oracle
    from pydream_b200 import targets
    out = str(tmp_path / 'r%d.pt')
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(2, port, out), nprocs=2, join=True)
    d, N, T, thin = 10, 8, 24, 3
    rng = np.random.default_rng(7)
    hist = rng.normal(size=(40, d))
    tgt = targets.BimodalMixture.benchmark(d)
    ref_s = c_oracle.OracleSampler(d, N, hist, hist[:N], tgt.kind, tgt.table(), seed=21, snooker=.2, history_thin=thin, multitry=3)
    ref = ref_s.run(T)
    r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    np.testing.assert_array_equal(np.concatenate([r0['dec'], r1['dec']], axis=1), ref['decisions'])
    np.testing.assert_array_equal(np.concatenate([r0['states'], r1['states']], axis=1), ref['states'])
    np.testing.assert_array_equal(np.concatenate([r0['logp'], r1['logp']], axis=1), ref['logp'])
    np.testing.assert_array_equal(r0['hist'], ref_s.history_flat)      # both replicas of the archive == the 1-process archive
    np.testing.assert_array_equal(r1['hist'], ref_s.history_flat)
