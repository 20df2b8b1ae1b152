"""pydream_b200: B200-native MT-DREAM(ZS) step path behind the PyDREAM API.

    from pydream_b200.core import run_dream
    from pydream_b200.parameters import SampledParam, FlatParam
    from pydream_b200.convergence import Gelman_Rubin
    from pydream_b200 import targets

Importing the package does not import torch; the engine does, and it fails loudly when no CUDA
device or no libdreamzs.so is present (there is no CPU fallback).
"""
__all__ = ['core', 'Dream', 'model', 'parameters', 'convergence', 'targets', 'engine', 'build']
