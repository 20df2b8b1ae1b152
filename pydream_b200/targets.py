"""Analytic benchmark log-likelihoods that the fused sm_100a step kernel evaluates in-register.

Each target is a plain callable ``likelihood(param_vec) -> float`` (so it can be handed to the
reference's ``run_dream`` unchanged) that also carries the constants the kernel needs.  The host
``__call__`` restates the arithmetic of the reference's example scripts:

* `CorrelatedGaussian`  pydream/examples/ndim_gaussian/dream_ex_ndim_gaussian.py:29-52
* `BimodalMixture`      pydream/examples/mixturemodel/mixturemodel.py:18-48
* `SumShift`            pydream/tests/test_models.py:46-50 (``simple_likelihood``)
* `Constant`            BASELINE.json configs[0] ("likelihood == 0")
* `Banana`              not in the reference: Haario et al. twisted Gaussian,
                        x2 <- x2 + b*x1^2 - 100*b, then N(0, diag(100, 1, ..., 1)) up to a constant.
"""
import numpy as np

# kernel-side target kinds (must match csrc/dreamzs_common.cuh and include/dreamzs.h)
TARGET_CONSTANT = 0
TARGET_GAUSSIAN_DENSE = 1
TARGET_MIXTURE = 2
TARGET_BANANA = 3
TARGET_SUMSHIFT = 4
TARGET_EXTERNAL = 5


class AnalyticTarget:
    """Base class: ``kind`` + flat float64 constant table understood by the kernel."""
    kind = None

    def table(self):
        raise NotImplementedError

    @property
    def ndim(self):
        raise NotImplementedError


class Constant(AnalyticTarget):
    kind = TARGET_CONSTANT

    def __init__(self, ndim, value=0.0):
        self._ndim = int(ndim)
        self.value = float(value)

    ndim = property(lambda self: self._ndim)

    def table(self):
        return np.array([self.value], dtype=np.float64)

    def __call__(self, param_vec):
        return self.value


class SumShift(AnalyticTarget):
    """``np.sum(param + shift)``."""
    kind = TARGET_SUMSHIFT

    def __init__(self, ndim, shift=3.0):
        self._ndim = int(ndim)
        self.shift = float(shift)

    ndim = property(lambda self: self._ndim)

    def table(self):
        return np.array([self.shift], dtype=np.float64)

    def __call__(self, param_vec):
        return np.sum(param_vec + self.shift)


class CorrelatedGaussian(AnalyticTarget):
    """``log_F - .5 * sum(x * dot(invC, x))`` with a dense precision matrix."""
    kind = TARGET_GAUSSIAN_DENSE

    def __init__(self, invC, log_F=0.0):
        self.invC = np.ascontiguousarray(invC, dtype=np.float64)
        assert self.invC.ndim == 2 and self.invC.shape[0] == self.invC.shape[1]
        self.log_F = float(log_F)

    @classmethod
    def benchmark(cls, d):
        """The covariance of the reference example: C_ij = (.5 delta_ij + .5) sqrt((i+1)(j+1))."""
        A = .5 * np.identity(d) + .5 * np.ones((d, d))
        idx = np.arange(1, d + 1, dtype=np.float64)
        C = A * np.sqrt(np.outer(idx, idx))
        invC = np.linalg.inv(C)
        if d > 150:
            log_F = 0.0
        else:
            log_F = np.log(((2 * np.pi) ** (-d / 2)) * np.linalg.det(C) ** (- 1. / 2))
        return cls(invC, log_F)

    ndim = property(lambda self: self.invC.shape[0])

    def table(self):
        return np.concatenate([[self.log_F], self.invC.reshape(-1)])

    def __call__(self, param_vec):
        return self.log_F - .5 * np.sum(param_vec * np.dot(self.invC, param_vec))


class BimodalMixture(AnalyticTarget):
    """Two unit-covariance Gaussians: ``logsumexp_j(-.5 |x - mu_j|^2 + log_F_j)``."""
    kind = TARGET_MIXTURE

    def __init__(self, mu, log_F):
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        self.log_F = np.ascontiguousarray(log_F, dtype=np.float64)
        assert self.mu.shape[0] == 2 and self.log_F.shape == (2,)

    @classmethod
    def benchmark(cls, d=10):
        mu = np.array([np.linspace(-5, -5, num=d), np.linspace(5, 5, num=d)])
        return cls(mu, np.array([-10.2880, -9.5949]))

    ndim = property(lambda self: self.mu.shape[1])

    def table(self):
        return np.concatenate([self.log_F, self.mu.reshape(-1)])

    def __call__(self, params):
        log_lh = np.zeros((2))
        for j in range(2):
            log_lh[j] = -.5 * np.sum((params - self.mu[j, :]) ** 2) + self.log_F[j]
        maxll = np.max(log_lh)
        density = np.sum(np.exp(log_lh - maxll))
        return np.log(density) + maxll


class Banana(AnalyticTarget):
    kind = TARGET_BANANA

    def __init__(self, ndim, b=0.1, var1=100.0):
        assert ndim >= 2
        self._ndim = int(ndim)
        self.b = float(b)
        self.var1 = float(var1)

    ndim = property(lambda self: self._ndim)

    def table(self):
        return np.array([self.b, self.var1], dtype=np.float64)

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        y2 = x[1] + self.b * (x[0] * x[0]) - self.var1 * self.b
        ss = (x[0] * x[0]) / self.var1 + y2 * y2
        if x.shape[0] > 2:
            ss = ss + np.sum(x[2:] * x[2:])
        return -.5 * ss


class TorchLikelihood(AnalyticTarget):
    """Escape hatch for likelihoods that are not one of the in-kernel analytic targets:
    ``fn(points)`` receives a float64 CUDA tensor ``[n, ndim]`` (one proposal per chain) and returns the ``n``
    log-likelihoods as a float64 CUDA tensor.  The step is then split into dreamzs_propose -> fn -> dreamzs_accept
    (one launch pair per iteration, no window fusion); everything else of the step stays on the device.
    As a host callable (``likelihood(param_vec) -> float``, pydream/model.py:30) it evaluates ``fn`` on one point."""
    kind = TARGET_EXTERNAL

    def __init__(self, ndim, fn):
        self._ndim = int(ndim)
        self.fn = fn

    ndim = property(lambda self: self._ndim)

    def table(self):
        return np.array([0.0], dtype=np.float64)

    def evaluate(self, points):
        out = self.fn(points)
        if out.dtype != points.dtype or out.shape != (points.shape[0],) or out.device != points.device:
            raise ValueError('TorchLikelihood.fn must return a float64 tensor of shape [n] on the device of its input')
        return out.contiguous()

    def __call__(self, param_vec):
        import torch
        x = torch.as_tensor(np.asarray(param_vec, dtype=np.float64).reshape(1, -1), device='cuda')
        return float(self.evaluate(x)[0].item())


class HostLikelihood(TorchLikelihood):
    """The reference's own likelihood contract, ``likelihood(param_vec: np.ndarray[ndim]) -> float`` (pydream/model.py:30):
    an arbitrary Python callable evaluated on the HOST, one point at a time (or by ``pool.map`` when a pool is given, the
    counterpart of Dream's ``parallel=True``).  The MT-DREAM(ZS) step itself stays on the GPU (dreamzs_propose /
    dreamzs_select / dreamzs_accept); only the proposals travel to the host and their log-likelihoods back, as the
    reference hands them to the user's function.  run_dream wraps plain callables in this class."""

    def __init__(self, ndim, fn, pool=None):
        super().__init__(ndim, fn)
        self.pool = pool

    def evaluate(self, points):
        import torch
        pts = points.detach().cpu().numpy()
        if self.pool is not None:
            vals = self.pool.map(self.fn, [p for p in pts])
        else:
            vals = [self.fn(p) for p in pts]
        out = np.asarray([float(v) for v in vals], dtype=np.float64)
        return torch.from_numpy(out).to(points.device)

    def __call__(self, param_vec):
        return float(self.fn(np.asarray(param_vec, dtype=np.float64)))
