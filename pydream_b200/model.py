"""``Model``: priors + likelihood of one sampled point, interface of pydream/model.py:8-32
(``Model(likelihood, sampled_parameters)``, ``total_logp(q0) -> (prior_logp, loglike)``).
Used on the host only (initial checks, user likelihoods); analytic targets are evaluated in-kernel."""
import numpy as np


class Model():

    def __init__(self, likelihood, sampled_parameters):
        self.likelihood = likelihood
        self.sampled_parameters = sampled_parameters if type(sampled_parameters) is list else [sampled_parameters]
        edges = np.concatenate([[0], np.cumsum([int(p.dsize) for p in self.sampled_parameters])])
        self._spans = [slice(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:])]

    def total_logp(self, q0):
        scalar_point = np.ndim(q0) == 0   # the reference falls back to the whole value for a scalar q0
        prior_logp = 0
        for param, span in zip(self.sampled_parameters, self._spans):
            prior_logp += param.prior(q0 if scalar_point else q0[span])
        return prior_logp, self.likelihood(q0)
