"""Gelman-Rubin diagnostic with the interface of pydream/convergence.py:3-20, computed by the
sm_100a reduction kernels (dreamzs_gr_chain_stats / dreamzs_gr_finish)."""
import numpy as np
import torch

from .engine import gelman_rubin_device, round_up4


def Gelman_Rubin(sampled_parameters):
    """sampled_parameters: list (one per chain) of (nsamples, ndim) arrays, or a device tensor
    [nchains, nsamples, ld] as returned by DreamEngine.run.  Returns Rhat[ndim] (numpy)."""
    if isinstance(sampled_parameters, torch.Tensor):
        return gelman_rubin_device(sampled_parameters, sampled_parameters.shape[2]).cpu().numpy()
    arr = np.stack([np.asarray(c, dtype=np.float64) for c in sampled_parameters])
    n, T_, d = arr.shape
    ld = round_up4(d)
    host = np.zeros((n, T_, ld))
    host[:, :, :d] = arr
    dev = torch.from_numpy(host).cuda()
    return gelman_rubin_device(dev, d).cpu().numpy()
