"""Gelman-Rubin diagnostic with the interface of pydream/convergence.py:3-20, computed by the
sm_100a reduction kernels (dreamzs_gr_chain_stats / dreamzs_gr_finish).  Like the rest of the path it has no
CPU implementation in the product: the CPU restatement lives under oracle/ and is test infrastructure."""
import numpy as np
import torch

from .engine import gelman_rubin_device, round_up4


def Gelman_Rubin(sampled_parameters, ndim=None):
    """sampled_parameters: list (one per chain) of (nsamples, ndim) arrays, or a device tensor
    [nchains, nsamples, ld] as returned by DreamEngine.run (rows padded to ld = ndim rounded up to 4: pass `ndim`,
    else every column of the tensor is treated as a dimension).  Returns Rhat[ndim] (numpy)."""
    if isinstance(sampled_parameters, torch.Tensor):
        d = int(ndim) if ndim is not None else sampled_parameters.shape[2]
        return gelman_rubin_device(sampled_parameters, d).cpu().numpy()
    arr = np.stack([np.asarray(c, dtype=np.float64) for c in sampled_parameters])
    n, T_, d = arr.shape
    ld = round_up4(d)
    host = np.zeros((n, T_, ld))
    host[:, :, :d] = arr
    dev = torch.from_numpy(host).cuda()
    return gelman_rubin_device(dev, d).cpu().numpy()
