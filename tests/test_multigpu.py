"""Sharded run (one process per GPU, NCCL archive all-gather + all-reduced adaptation) must reproduce the
single-GPU trajectories bit for bit: chain ids in the Philox counter are global and the archive layout
is shard-independent.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    'gauss100': dict(d=100, N=64, T=45, target='gaussian', kw=dict(snooker=.1, history_thin=10)),
    'gauss50_adapt': dict(d=50, N=48, T=60, target='gaussian',
                          kw=dict(snooker=.1, history_thin=5, adapt_crossover=True, crossover_burnin=40)),
    'banana200': dict(d=200, N=32, T=25, target='banana', kw=dict(snooker=.1, history_thin=5)),
}


def _inputs(case):
    from golden_util import make_target
    rng = np.random.default_rng(123)
    d, N = case['d'], case['N']
    hist = rng.uniform(-5, 15, size=(2 * N + 9, d))
    return make_target(dict(kind=case['target'], d=d)), hist, hist[:N].copy()


def _worker(rank, world, port, name, out, peer_archive):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from pydream_b200.engine import DreamEngine
    case = CASES[name]
    tgt, hist, starts = _inputs(case)
    eng = DreamEngine(case['d'], case['N'], hist, starts, tgt, seed=4, group=dist.group.WORLD, peer_archive=peer_archive,
                      **case['kw'])
    # two calls: the archive grows between them (collective re-allocation of the shared block)
    T1 = case['T'] // 3
    parts = [eng.run(T1), eng.run(case['T'] - T1)]
    trace, logp, dec = (torch.cat([a[i] for a in parts], dim=1) for i in range(3))
    rhat = eng.gelman_rubin(trace)
    torch.cuda.synchronize()
    eng.check_peers()
    torch.save(dict(trace=trace.cpu(), logp=logp.cpu(), dec=dec.cpu(), Z=eng.Z[:eng.archive_rows].cpu(),
                    cr=eng.cr_probs.cpu(), rhat=rhat.cpu(), peers=eng.peers is not None), out % rank)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('peer_archive', [True, False], ids=['nvlink_peer_stores', 'nccl_allgather'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_sharded_equals_single_gpu(name, peer_archive, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from pydream_b200.engine import DreamEngine
    case = CASES[name]
    tgt, hist, starts = _inputs(case)
    eng = DreamEngine(case['d'], case['N'], hist, starts, tgt, seed=4, **case['kw'])
    trace, logp, dec = eng.run(case['T'])
    rhat = eng.gelman_rubin(trace).cpu()
    out = str(tmp_path / 'rank%d.pt')
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, name, out, peer_archive), nprocs=2, join=True)
    parts = [torch.load(out % r) for r in range(2)]
    assert torch.equal(torch.cat([p['trace'] for p in parts]), trace.cpu())
    assert torch.equal(torch.cat([p['logp'] for p in parts]), logp.cpu())
    assert torch.equal(torch.cat([p['dec'] for p in parts]), dec.cpu())
    for p in parts:   # every rank holds the full archive, identical to the single-GPU one
        assert p['peers'] == peer_archive
        assert torch.equal(p['Z'], eng.Z[:eng.archive_rows].cpu())
        np.testing.assert_allclose(p['cr'].numpy(), eng.cr_probs.cpu().numpy(), rtol=1e-12)
        np.testing.assert_allclose(p['rhat'].numpy(), rhat.numpy(), rtol=1e-12)


def _worker_run_dream(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from scipy.stats import norm
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from pydream_b200 import targets
    from pydream_b200.core import run_dream
    from pydream_b200.parameters import SampledParam
    params = [SampledParam(norm, loc=np.array([-6.6, 3, 1.0, -.12]), scale=np.array([.13, 5, .9, 1.0]))]
    s, l = run_dream(params, targets.SumShift(4, 3.0), niterations=50, nchains=8, verbose=False, save_history=False,
                     seed=1234, group=dist.group.WORLD)
    np.savez(out % rank, s=np.stack(s), l=np.stack(l))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_run_dream_draws_on_rank0(tmp_path):
    """run_dream(group=...) WITHOUT history_file / start: the archive seed, the random starts (prior draws on the host) and
    the seed are drawn on rank 0 and broadcast, so the sharded run equals the single-GPU run with the same seed."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from scipy.stats import norm
    from pydream_b200 import targets
    from pydream_b200.core import run_dream
    from pydream_b200.parameters import SampledParam
    params = [SampledParam(norm, loc=np.array([-6.6, 3, 1.0, -.12]), scale=np.array([.13, 5, .9, 1.0]))]
    s1, l1 = run_dream(params, targets.SumShift(4, 3.0), niterations=50, nchains=8, verbose=False, save_history=False, seed=1234)
    out = str(tmp_path / 'rd%d.npz')
    mp.spawn(_worker_run_dream, args=(2, 29700 + os.getpid() % 1000, out), nprocs=2, join=True)
    parts = [np.load(out % r) for r in range(2)]
    np.testing.assert_array_equal(np.concatenate([p['s'] for p in parts]), np.stack(s1))
    np.testing.assert_array_equal(np.concatenate([p['l'] for p in parts]), np.stack(l1))
