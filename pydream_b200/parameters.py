"""Parameter priors with the interface of pydream/parameters.py (SampledParam :6-47, FlatParam :49-70).

``SampledParam`` wraps a frozen SciPy distribution exactly as the reference does; the GPU step
evaluates ``scipy.stats.norm`` and ``scipy.stats.uniform`` priors in closed form in-kernel and
``FlatParam`` as log prior 0 with infinite bounds.
"""
import numpy as np


class SampledParam():
    """A SciPy-based parameter prior class (same constructor and methods as the reference)."""

    def __init__(self, scipy_distribution, *args, **kwargs):
        self.dist = scipy_distribution(*args, **kwargs)
        self.dsize = self.random().size

    def interval(self, alpha=1):
        return self.dist.interval(alpha)

    def random(self, reseed=False, random_state=None):
        """One draw (pydream/parameters.py:27-35).  `random_state` (extension): a numpy Generator / RandomState to draw
        from, which is how run_dream(seed=...) makes its prior draws reproducible."""
        if random_state is None:
            random_state = np.random.RandomState() if reseed else None
        return self.dist.rvs(random_state=random_state)

    def prior(self, q0):
        return np.sum(self.dist.logpdf(q0))

    def closed_form(self):
        """-> (kind, loc[dsize], scale[dsize]) in the encoding of include/dreamzs.h, or None when the
        distribution has no in-kernel closed form."""
        name = getattr(getattr(self.dist, 'dist', None), 'name', None)
        if name not in ('norm', 'uniform') or self.dist.args[2:] or set(self.dist.kwds) - {'loc', 'scale'}:
            return None
        args = list(self.dist.args)
        loc = self.dist.kwds.get('loc', args[0] if len(args) > 0 else 0.0)
        scale = self.dist.kwds.get('scale', args[1] if len(args) > 1 else 1.0)
        loc = np.broadcast_to(np.asarray(loc, dtype=np.float64), (self.dsize,)).copy()
        scale = np.broadcast_to(np.asarray(scale, dtype=np.float64), (self.dsize,)).copy()
        return (1 if name == 'norm' else 2), loc, scale


class FlatParam(SampledParam):
    """A Flat parameter class (returns 0 at all locations); ``test_value`` fixes the dimension."""

    def __init__(self, test_value):
        self.dsize = test_value.size

    def prior(self, q0):
        return 0

    def interval(self, alpha=1):
        lower = [-np.inf] * self.dsize
        upper = [np.inf] * self.dsize
        return [lower, upper]

    def closed_form(self):
        return 0, np.zeros(self.dsize), np.ones(self.dsize)
