"""``run_dream``: drop-in for pydream.core.run_dream (pydream/core.py:11-86) on one or more B200s.

Same signature, keyword names, defaults, checks, exception texts and return shapes as the
reference.  The multiprocessing pool (core.py:250-327) is replaced by `engine.DreamEngine`: all
chains advance in lock-step inside one fused sm_100a kernel.  Extra, optional keywords (not in the
reference): ``seed`` (Philox key; default drawn from the OS like the reference's unseeded runs),
``device``, ``group`` (torch.distributed process group: chains are sharded over its ranks and each
rank returns the chains it owns), ``return_device`` (return device tensors instead of numpy lists), ``stream_chunk`` (iterations per
device->host chunk; samples stream to pinned host memory while sampling continues).
"""
import os
from datetime import datetime

import numpy as np

from .Dream import Dream
from .model import Model
from . import targets as T


def _prior_arrays(parameters):
    kinds, a, b = [], [], []
    for p in parameters:
        cf = p.closed_form() if hasattr(p, 'closed_form') else None
        if cf is None:
            raise NotImplementedError(
                'pydream_b200 evaluates FlatParam, scipy.stats.norm and scipy.stats.uniform priors in-kernel; '
                'prior %r has no closed form here' % (p,))
        k, loc, scale = cf
        kinds.append(np.full(p.dsize, k, dtype=np.int32))
        a.append(loc)
        b.append(scale)
    return np.concatenate(kinds), np.concatenate(a), np.concatenate(b)


def _has_history(step):
    hf = step.history_file
    return isinstance(hf, np.ndarray) or hf != False   # noqa: E712


def _reference_archive_rows(nchains, niterations, step, len_old_history):
    """Archive sizing of _setup_mp_dream_pool (pydream/core.py:260-268), in rows."""
    d, thin = step.total_var_dimension, step.history_thin
    seed_len = len_old_history if _has_history(step) else step.nseedchains * d
    if niterations < thin:
        arr_dim = ((np.floor(nchains*niterations/thin)+nchains)*d)+seed_len
    else:
        arr_dim = np.floor(((nchains*niterations*d)/thin))+seed_len
    return int(arr_dim) // d


def run_dream(parameters, likelihood, nchains=5, niterations=50000, start=None, restart=False, verbose=True,
              nverbose=10, tempering=False, mp_context=None, **kwargs):
    """Run MT-DREAM(ZS); returns (sampled_params, log_ps): per chain an (niterations, ndim) array and an
    (niterations, 1) array, as pydream.core.run_dream does."""
    seed = kwargs.pop('seed', None)
    device = kwargs.pop('device', None)
    group = kwargs.pop('group', None)
    return_device = kwargs.pop('return_device', False)
    stream_chunk = kwargs.pop('stream_chunk', 256)

    if restart:
        if start == None:   # noqa: E711
            raise Exception('Restart run specified but no start positions given.')
        if 'model_name' not in kwargs:
            raise Exception('Restart run specified but no model name to load history and crossover value files from given.')
    if type(parameters) is not list:
        parameters = [parameters]
    model = Model(likelihood=likelihood, sampled_parameters=parameters)
    if restart:
        step_instance = Dream(model=model, variables=parameters,
                              history_file=kwargs['model_name'] + '_DREAM_chain_history.npy',
                              crossover_file=kwargs['model_name'] + '_DREAM_chain_adapted_crossoverprob.npy',
                              gamma_file=kwargs['model_name'] + '_DREAM_chain_adapted_gammalevelprob.npy',
                              verbose=verbose, mp_context=mp_context, **kwargs)
    else:
        step_instance = Dream(model=model, variables=parameters, verbose=verbose, mp_context=mp_context, **kwargs)

    d = step_instance.total_var_dimension
    # ---- checks and sizing of _setup_mp_dream_pool (pydream/core.py:250-305), same messages
    min_njobs = (2*len(step_instance.DEpairs))+1
    if nchains < min_njobs:
        raise Exception('Dream should be run with at least (2*DEpairs)+1 number of chains.  For current algorithmic settings, set njobs>=%s.' % str(min_njobs))
    old_history = None
    len_old_history = 0
    if _has_history(step_instance):
        # a path, as in the reference, or (extension) the array itself
        hf = step_instance.history_file
        old_history = hf if isinstance(hf, np.ndarray) else np.load(hf)
        len_old_history = int(np.asarray(old_history).size)
        step_instance.nseedchains = len_old_history/d
    min_nseedchains = 2*len(step_instance.DEpairs)*nchains
    if step_instance.nseedchains < min_nseedchains:
        raise Exception('The size of the seeded starting history is insufficient.  Increase nseedchains>=%s.' % str(min_nseedchains))
    if step_instance.crossover_burnin == None:   # noqa: E711
        step_instance.crossover_burnin = int(np.floor(niterations/10))
    if start is not None:
        if step_instance.start_random:
            print('Warning: start position provided but random_start set to True.  Overrode random_start value and starting walk at provided start position.')
            step_instance.start_random = False

    if not isinstance(likelihood, T.AnalyticTarget):
        # the reference's contract, likelihood(param_vec) -> float (pydream/model.py:30): the step stays on the GPU
        # (split into propose / select / accept), the user's function is called on the host for every proposal
        if not callable(likelihood):
            raise TypeError('likelihood must be a callable or a target from pydream_b200.targets')
        likelihood = T.HostLikelihood(d, likelihood)
    if likelihood.ndim != d:
        raise ValueError('target dimension %d != total parameter dimension %d' % (likelihood.ndim, d))
    prior_kind, prior_a, prior_b = _prior_arrays(parameters)

    # ---- seed of the run.  Everything random on the host side (archive seed, random starts) is drawn from a numpy
    #      Philox generator keyed by it, the device side from the Philox contract of DESIGN.md keyed by it: a run with
    #      `seed=` given is reproducible.  Sharded runs draw on rank 0 and broadcast, so every rank seeds the same
    #      archive replica and the same chains as the single-GPU run.
    sharded = group is not None and _world_size(group) > 1
    if seed is None:
        seed = int.from_bytes(os.urandom(8), 'little')
        if sharded:
            seed = _broadcast_from_rank0(seed, group)
    seed = int(seed) & (2 ** 64 - 1)
    host_rng = np.random.Generator(np.random.Philox(key=[seed, 0x5EED]))

    # ---- archive seed (Dream.py:203-214) and start positions (core.py:74-78, Dream.py:221-225)
    drawn = None
    if old_history is not None:
        history = np.asarray(old_history, dtype=np.float64).reshape(-1, d)
    else:
        nseed = int(step_instance.nseedchains)
        if not sharded or _rank(group) == 0:
            drawn = np.array([step_instance.draw_from_prior(step_instance.variables, rng=host_rng) for _ in range(nseed)]).reshape(nseed, d)
        history = _broadcast_from_rank0(drawn, group) if sharded else drawn
    if step_instance.start_random:
        drawn = None
        if not sharded or _rank(group) == 0:
            drawn = np.array([step_instance.draw_from_prior(step_instance.variables, random_seed=True, rng=host_rng) for _ in range(nchains)])
        starts = _broadcast_from_rank0(drawn, group) if sharded else drawn
    elif type(start) is list:
        starts = np.array([np.asarray(s, dtype=np.float64).reshape(-1) for s in start[:nchains]])
    else:
        starts = np.tile(np.asarray(start, dtype=np.float64).reshape(1, -1), (nchains, 1))
    starts = starts.reshape(nchains, d)

    import time
    timing = os.environ.get('DREAMZS_TIMING')
    t_a = time.perf_counter()
    from .engine import DreamEngine   # imports torch; fails loudly without CUDA / libdreamzs.so
    eng = DreamEngine(d, nchains, history, starts, likelihood, prior_kind, prior_a, prior_b, seed=seed,
                      nCR=step_instance.nCR, gamma_levels=step_instance.ngamma, DEpairs=len(step_instance.DEpairs),
                      multitry=int(step_instance.multitry), snooker=step_instance.snooker,
                      p_gamma_unity=step_instance.p_gamma_unity, lamb=step_instance.lamb, zeta=step_instance.zeta,
                      history_thin=step_instance.history_thin, hardboundaries=bool(step_instance.boundaries),
                      adapt_crossover=step_instance.adapt_crossover, adapt_gamma=step_instance.adapt_gamma,
                      crossover_burnin=step_instance.crossover_burnin,
                      cr_probs=np.asarray(step_instance.CR_probabilities, dtype=np.float64),
                      gamma_probs=np.asarray(step_instance.gamma_probabilities, dtype=np.float64),
                      device=device, group=group, record_decisions=bool(verbose), reserve_iters=niterations)
    import torch
    t_b = time.perf_counter()
    acc_chunks = []
    on_chunk = None
    if verbose:
        on_chunk = lambda dec, t0: acc_chunks.append((dec & 1).to(torch.float64).mean(dim=0))
    if tempering:
        # _sample_dream_pt (pydream/core.py:131-236): two records per iteration (after the step, after the swap)
        trace, logp, dec, swaps = eng.run_tempered(niterations)
        if verbose:
            _print_tempering(dec, swaps, niterations, nchains)
    elif return_device:
        trace, logp, dec = eng.run(niterations)
        if verbose and dec is not None:
            on_chunk(dec, 0)
    else:
        # samples stream to pinned host memory while sampling continues (torch's caching host allocator
        # recycles the pinned blocks of earlier calls once their arrays have been dropped)
        tr_host = torch.empty((eng.Nl, niterations, d), dtype=torch.float64, pin_memory=True)
        lp_host = torch.empty((eng.Nl, niterations, 1), dtype=torch.float64, pin_memory=True)
        eng.run_to_host(niterations, tr_host, lp_host, chunk_iters=stream_chunk, on_chunk=on_chunk)
    t_c = time.perf_counter()
    step_instance.CR_probabilities = eng.cr_probs.cpu().numpy()
    step_instance.gamma_probabilities = eng.gamma_probs.cpu().numpy()

    if verbose and acc_chunks:
        _print_acceptance(torch.cat(acc_chunks), niterations, nverbose)

    # ---- history / adapted probabilities on disk when the archive is full (Dream.py:939-969)
    if step_instance.save_history and eng.rank == 0:
        ref_rows = _reference_archive_rows(nchains, niterations, step_instance, len_old_history)
        if eng.archive_rows >= ref_rows:
            prefix = (step_instance.model_name + '_') if step_instance.model_name else datetime.now().strftime('%Y_%m_%d_%H:%M:%S') + '_'
            step_instance.save_history_to_disc(eng.history_flat()[:ref_rows * d], prefix)

    if return_device:
        eng.check_peers()
        return trace, logp     # (a sharded caller keeps `trace` alive; the shared archive is released with the engine)
    if tempering:
        # arrays (nchains, 2 niterations, ndim) and (nchains, 2 niterations, 1), as core.py:145-146, 236
        sampled_params = trace[:, :, :d].contiguous().cpu().numpy()
        log_ps = logp.unsqueeze(2).cpu().numpy()
        eng.close()
        return sampled_params, log_ps
    torch.cuda.current_stream(eng.device).synchronize()
    eng.check_peers()
    eng.close()
    t_d = time.perf_counter()
    tr_np, lp_np = tr_host.numpy(), lp_host.numpy()
    sampled_params = [tr_np[c] for c in range(eng.Nl)]
    log_ps = [lp_np[c] for c in range(eng.Nl)]
    if timing:
        print('run_dream stages (ms): engine+upload %.2f, enqueue %.2f, drain %.2f, lists %.2f'
              % (1e3 * (t_b - t_a), 1e3 * (t_c - t_b), 1e3 * (t_d - t_c), 1e3 * (time.perf_counter() - t_d)))
    return sampled_params, log_ps


def _world_size(group):
    import torch.distributed as dist
    return dist.get_world_size(group)


def _rank(group):
    import torch.distributed as dist
    return dist.get_rank(group)


def _broadcast_from_rank0(obj, group):
    """The object rank 0 of `group` holds, on every rank (host-side inputs of a sharded run_dream)."""
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0), group=group)
    return box[0]


def _print_tempering(dec, swaps, niterations, nchains):
    """The progress lines of _sample_dream_pt (pydream/core.py:159-171) every 10 iterations, from the decision words:
    per-chain acceptance rate of the steps and the rate of accepted temperature swaps."""
    import torch
    acc = torch.cumsum((dec[:, 0::2] & 1).to(torch.float64), dim=1).cpu().numpy()          # [N, niter]
    sw = torch.cumsum(swaps[:, 2], dim=0).cpu().numpy()
    for iteration in range(0, niterations, 10):
        naccepts = acc[:, iteration - 1] if iteration > 0 else np.zeros(nchains)
        nswaps = sw[iteration - 1] if iteration > 0 else 0.0
        print('Iteration: ', iteration, ' overall acceptance rate: ', naccepts/(iteration/float(nchains) + iteration + 1),
              ' and overall temp swap acceptance rate: ', nswaps/(iteration+1))


def _print_acceptance(acc_mean, niterations, nverbose):
    """Acceptance-rate lines at the cadence of _sample_dream (pydream/core.py:104-112), averaged over chains
    (the reference prints one line per chain process).  acc_mean: [T] mean acceptance per iteration (device)."""
    import torch
    cum = torch.cumsum(acc_mean, dim=0).cpu().numpy()
    for iteration in range(0, niterations, max(int(nverbose), 1)):
        naccepts = cum[iteration - 1] if iteration > 0 else 0.0
        print('Iteration: ', iteration, ' acceptance rate: ', float(naccepts)/(iteration+1))
        if iteration % 100 == 0:
            lo = cum[iteration - 101] if iteration > 100 else 0.0
            win = (naccepts - lo) if iteration > 0 else 0.0
            print('Iteration: ', iteration, ' acceptance rate over last 100 iterations: ', float(win)/100)
