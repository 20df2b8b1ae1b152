"""Random option combinations (tools/fuzz_oracle.py's generator: both adaptations, 1-4 gamma levels, every closed-form
prior, no hard boundaries, multi-try 3-5, DE pairs 1-3, zeta 1e-3, parallel tempering) through the CUDA path and the C
oracle: decision words identical, log-posteriors within 10x the tolerance, states rtol 1e-9.  The oracle itself is
pinned to the unmodified reference on the same generator (tests/test_oracle_golden.py, tools/fuzz_oracle.py)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))


@pytest.mark.parametrize('i', range(48))
def test_cuda_matches_oracle_on_random_options(i):
    import fuzz_cuda
    meta, hist = fuzz_cuda.random_case(i)
    fuzz_cuda.check(meta, hist)
