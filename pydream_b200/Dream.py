"""``Dream``: the sampler options of pydream/Dream.py:63-191 (same keyword names, defaults, warnings
and derived attributes).  In this framework the object is a validated option set; the step itself
(``astep``, Dream.py:193-422) runs for all chains at once inside the fused sm_100a kernel driven by
``pydream_b200.engine.DreamEngine``.
"""
import numpy as np


class Dream():

    def __init__(self, model, variables=None, nseedchains=None, nCR=3, adapt_crossover=True, adapt_gamma=False,
                 crossover_burnin=None, DEpairs=1, lamb=.05, zeta=1e-12, history_thin=10, snooker=.10,
                 p_gamma_unity=.20, gamma_levels=1, start_random=True, save_history=True, history_file=False,
                 crossover_file=False, gamma_file=False, multitry=False, parallel=False, verbose=False,
                 model_name=False, hardboundaries=True, mp_context=None, **kwargs):
        self.mp_context = mp_context
        self.model = model
        self.model_name = model_name
        self.variables = self.model.sampled_parameters if variables is None else variables
        self.boundaries = hardboundaries
        self.total_var_dimension = 0
        for var in self.variables:
            self.total_var_dimension += var.dsize
        if self.boundaries:
            self.boundary_mask = True if self.total_var_dimension == 1 else np.ones((self.total_var_dimension), dtype=bool)
            self.mins, self.maxs = [], []
            for var in self.variables:
                interval = var.interval(1)
                if var.dsize > 1:
                    self.mins += list(interval[0])
                    self.maxs += list(interval[1])
                else:
                    self.mins.append(interval[0])
                    self.maxs.append(interval[1])
            self.mins = np.array(self.mins, dtype=np.float64).reshape(-1)
            self.maxs = np.array(self.maxs, dtype=np.float64).reshape(-1)
        self.nseedchains = nseedchains
        self.nCR = nCR
        if self.nCR > self.total_var_dimension:
            self.nCR = self.total_var_dimension
            print('Warning: the total number of crossover values specified ('+str(nCR)+') is less than the total dimension of all variables ('+str(self.total_var_dimension)+').  Setting the number of crossover values to be equal to the total variable dimension.')
        if self.total_var_dimension == 1 and adapt_crossover:
            adapt_crossover = False
            print('Warning: the total variable dimension = 1, so crossover values will not be adapted, even though crossover adaptation was requested.')
        self.ngamma = gamma_levels
        self.njoint_cr_gamma_probs = nCR*gamma_levels
        self.crossover_burnin = crossover_burnin
        self.crossover_file = crossover_file
        self.adapt_crossover = adapt_crossover
        if crossover_file:
            self.CR_probabilities = np.load(crossover_file)
            self.nCR = len(self.CR_probabilities)
            if self.adapt_crossover:
                print('Warning: Crossover values loaded and adapt_crossover = True.  Crossover values will be further adapted.')
        else:
            self.CR_probabilities = [1/float(self.nCR) for i in range(self.nCR)]
        self.adapt_gamma = adapt_gamma
        if gamma_file:
            self.gamma_probabilities = np.load(gamma_file)
            if adapt_gamma:
                print('Warning: Gamma values loaded and adapt gamma = True.  Gamma values will be further adapted.')
        else:
            self.gamma_probabilities = [1/float(self.ngamma) for i in range(self.ngamma)]
        self.CR_values = np.array([m/float(self.nCR) for m in range(1, self.nCR+1)])
        self.gamma_level_values = np.array([m for m in range(1, self.ngamma+1)])
        self.DEpairs = np.linspace(1, DEpairs, num=DEpairs, dtype=int)
        self.snooker = snooker
        self.p_gamma_unity = p_gamma_unity
        if multitry == False:   # noqa: E712  (same truthiness rules as the reference: 0/False -> 1, True/1 -> 5)
            self.multitry = 1
        elif multitry == True:  # noqa: E712
            self.multitry = 5
        else:
            self.multitry = multitry
        self.parallel = parallel
        self.lamb = lamb
        self.zeta = zeta
        self.last_logp = None
        if self.nseedchains == None:   # noqa: E711
            self.nseedchains = self.total_var_dimension*10
        from .engine import gamma_table
        self.gamma_arr = gamma_table(self.ngamma, DEpairs, self.total_var_dimension)
        self.gamma = None
        self.iter = 0
        self.chain_n = None
        self.nchains = None
        self.len_history = 0
        self.save_history = save_history
        self.history_file = history_file
        self.history_thin = history_thin
        self.start_random = start_random
        self.verbose = verbose
        self.logp = self.model.total_logp
        self.extra_kwargs = kwargs

    def draw_from_prior(self, model_vars, random_seed=False):
        """Draw from the priors (pydream/Dream.py:628-644), same exception text."""
        draw = np.array([])
        for variable in model_vars:
            try:
                var_draw = variable.random(reseed=random_seed)
            except AttributeError:
                raise Exception('Random draw from distribution for variable %s not implemented yet.' % variable)
            draw = np.append(draw, var_draw)
        return draw.flatten()

    def astep(self, q0, T=1., last_loglike=None, last_logprior=None):
        raise Exception('Dream should be run with multiple chains in parallel.  Set nchains > 1.  '
                        '(pydream_b200 steps all chains at once on the GPU: use pydream_b200.core.run_dream, '
                        'pydream_b200.engine.DreamEngine.run, or DreamEngine.astep -- the same operator contract '
                        'for every chain at once)')

    def save_history_to_disc(self, history, prefix):
        """Same three files and messages as pydream/Dream.py:947-969."""
        filename = prefix+'DREAM_chain_history.npy'
        print('Saving history to file: ', filename)
        np.save(filename, history)
        filename = prefix+'DREAM_chain_adapted_crossoverprob.npy'
        print('Saving fitted crossover values: ', self.CR_probabilities, ' to file: ', filename)
        np.save(filename, self.CR_probabilities)
        filename = prefix+'DREAM_chain_adapted_gammalevelprob.npy'
        print('Saving fitted gamma level values: ', self.gamma_probabilities, ' to file: ', filename)
        np.save(filename, self.gamma_probabilities)
