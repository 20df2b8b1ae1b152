"""``Dream``: the validated option set of one MT-DREAM(ZS) run.

The reference's ``Dream`` (pydream/Dream.py:63-191) is both the option holder and the per-process stepper.  Here the
stepping happens for all chains at once inside the sm_100a kernels (``pydream_b200.engine.DreamEngine``); this class
keeps the reference's keyword names, defaults, derived attributes and warning texts so that ``run_dream`` and user code
that inspects the object behave the same.  Option table below: name -> (default, reference line).
"""
import numpy as np

# Scalar options stored verbatim on the instance (name, default, pydream/Dream.py line of the assignment)
_PLAIN_OPTIONS = (
    ('snooker', .10, 152), ('p_gamma_unity', .20, 153), ('parallel', False, 162), ('lamb', .05, 164), ('zeta', 1e-12, 165),
    ('save_history', True, 186), ('history_file', False, 187), ('history_thin', 10, 188), ('start_random', True, 189),
    ('verbose', False, 190), ('crossover_burnin', None, 116), ('crossover_file', False, 117), ('model_name', False, 78),
    ('mp_context', None, 76), ('nseedchains', None, 107),
)
_WARN_NCR = ('Warning: the total number of crossover values specified (%s) is less than the total dimension of all variables '
             '(%s).  Setting the number of crossover values to be equal to the total variable dimension.')
_WARN_1D = ('Warning: the total variable dimension = 1, so crossover values will not be adapted, even though crossover '
            'adaptation was requested.')
_WARN_CR_FILE = 'Warning: Crossover values loaded and adapt_crossover = True.  Crossover values will be further adapted.'
_WARN_GAMMA_FILE = 'Warning: Gamma values loaded and adapt gamma = True.  Gamma values will be further adapted.'


def _bounds(variables):
    """Per-dimension support of the priors: ``var.interval(1)`` of every parameter, flattened (Dream.py:86-105)."""
    lo, hi = [], []
    for var in variables:
        a, b = var.interval(1)
        lo.append(np.atleast_1d(np.asarray(a, dtype=np.float64)).reshape(-1))
        hi.append(np.atleast_1d(np.asarray(b, dtype=np.float64)).reshape(-1))
    return np.concatenate(lo), np.concatenate(hi)


def _uniform_probs(n):
    return [1 / float(n)] * n


class Dream():

    def __init__(self, model, variables=None, nseedchains=None, nCR=3, adapt_crossover=True, adapt_gamma=False,
                 crossover_burnin=None, DEpairs=1, lamb=.05, zeta=1e-12, history_thin=10, snooker=.10,
                 p_gamma_unity=.20, gamma_levels=1, start_random=True, save_history=True, history_file=False,
                 crossover_file=False, gamma_file=False, multitry=False, parallel=False, verbose=False,
                 model_name=False, hardboundaries=True, mp_context=None, **kwargs):
        given = locals()
        for name, _default, _line in _PLAIN_OPTIONS:
            setattr(self, name, given[name])
        self.model = model
        self.logp = model.total_logp
        self.variables = model.sampled_parameters if variables is None else variables
        self.extra_kwargs = kwargs                       # unknown keywords are tolerated, as in the reference (:67)

        # ---- dimension and support (Dream.py:80-105)
        d = sum(var.dsize for var in self.variables)
        self.total_var_dimension = d
        self.boundaries = hardboundaries
        if hardboundaries:
            self.boundary_mask = True if d == 1 else np.ones(d, dtype=bool)
            self.mins, self.maxs = _bounds(self.variables)

        # ---- crossover values and their probabilities (Dream.py:107-146)
        self.nCR = nCR
        if nCR > d:
            self.nCR = d
            print(_WARN_NCR % (str(nCR), str(d)))
        if d == 1 and adapt_crossover:
            adapt_crossover = False
            print(_WARN_1D)
        self.adapt_crossover = adapt_crossover
        if crossover_file:
            self.CR_probabilities = np.load(crossover_file)
            self.nCR = len(self.CR_probabilities)
            if adapt_crossover:
                print(_WARN_CR_FILE)
        else:
            self.CR_probabilities = _uniform_probs(self.nCR)
        self.CR_values = np.arange(1, self.nCR + 1) / float(self.nCR)

        # ---- gamma levels (Dream.py:120-149, 173-179)
        self.ngamma = gamma_levels
        self.njoint_cr_gamma_probs = nCR * gamma_levels
        self.adapt_gamma = adapt_gamma
        if gamma_file:
            self.gamma_probabilities = np.load(gamma_file)
            if adapt_gamma:
                print(_WARN_GAMMA_FILE)
        else:
            self.gamma_probabilities = _uniform_probs(gamma_levels)
        self.gamma_level_values = np.arange(1, gamma_levels + 1)
        self.DEpairs = np.arange(1, DEpairs + 1, dtype=int)
        from .engine import gamma_table
        self.gamma_arr = gamma_table(gamma_levels, DEpairs, d)

        # ---- multi-try: False/0 -> 1 try, True/1 -> 5 tries, any other number is taken as given (Dream.py:155-161;
        #      `1 == True` in Python, so multitry=1 means five tries in the reference as well)
        self.multitry = 1 if multitry == False else (5 if multitry == True else multitry)   # noqa: E712

        if self.nseedchains is None:
            self.nseedchains = 10 * d                    # Dream.py:170-171
        # per-process stepping state of the reference object; kept for code that inspects it
        self.last_logp = self.gamma = self.chain_n = self.nchains = None
        self.iter = self.len_history = 0

    def draw_from_prior(self, model_vars, random_seed=False, rng=None):
        """One draw from every prior, concatenated (pydream/Dream.py:628-644); same exception text.  `rng` (extension):
        the numpy Generator run_dream derives from its `seed` keyword."""
        parts = []
        for variable in model_vars:
            try:      # FlatParam has no distribution to draw from: its `random` fails on the missing attribute
                value = variable.random(reseed=random_seed) if rng is None else variable.random(random_state=rng)
            except AttributeError:
                raise Exception('Random draw from distribution for variable %s not implemented yet.' % variable)
            parts.append(np.asarray(value, dtype=np.float64).reshape(-1))
        return np.concatenate(parts) if parts else np.array([])

    def astep(self, q0, T=1., last_loglike=None, last_logprior=None):
        raise Exception('Dream should be run with multiple chains in parallel.  Set nchains > 1.  '
                        '(pydream_b200 steps all chains at once on the GPU: use pydream_b200.core.run_dream, '
                        'pydream_b200.engine.DreamEngine.run, or DreamEngine.astep -- the same operator contract '
                        'for every chain at once)')

    def save_history_to_disc(self, history, prefix):
        """The three files of pydream/Dream.py:947-969 (archive, crossover and gamma-level probabilities)."""
        for stem, what, values in (('DREAM_chain_history.npy', 'Saving history to file: ', history),
                                   ('DREAM_chain_adapted_crossoverprob.npy', 'Saving fitted crossover values: ', self.CR_probabilities),
                                   ('DREAM_chain_adapted_gammalevelprob.npy', 'Saving fitted gamma level values: ', self.gamma_probabilities)):
            filename = prefix + stem
            if values is history:
                print(what, filename)
            else:
                print(what, values, ' to file: ', filename)
            np.save(filename, values)
