"""Random option combinations (tools/fuzz_oracle.py's generator: both adaptations, 1-4 gamma levels, every closed-form
prior, no hard boundaries, multi-try 3-5, DE pairs 1-3, zeta 1e-3, parallel tempering) through the CUDA path and the C
oracle: decision words identical, log-posteriors within 10x the tolerance, states rtol 1e-9.  The oracle itself is
pinned to the unmodified reference on the same generator (tests/test_oracle_golden.py, tools/fuzz_oracle.py).

Case 53 (sum-shift target, d = 2, 70 % snooker steps, three DE pairs, zeta 1e-3) is the one case of the first 80 that
leaves those bounds with IDENTICAL decisions: every snooker step projects the state on a line through an archive row,
and the 1-ulp difference between the kernel's fused / butterfly dot products and the oracle's sequential ones (the
reference itself uses BLAS there, a third order) grows ~4x per snooker step -- 1.6e-10 relative after 35 iterations.
It is kept in the suite twice: with the bound the drift reaches (decisions exact, states 1e-7), and as a strict xfail
at the common bound so that a change which removes -- or hides -- the drift is noticed."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))

DRIFT_CASE = 53


@pytest.mark.parametrize('i', [i for i in range(80) if i != DRIFT_CASE])
def test_cuda_matches_oracle_on_random_options(i):
    import fuzz_cuda
    meta, hist = fuzz_cuda.random_case(i)
    fuzz_cuda.check(meta, hist)


def test_snooker_drift_case_decisions_exact_states_bounded():
    import fuzz_cuda
    meta, hist = fuzz_cuda.random_case(DRIFT_CASE)
    assert meta['kw']['snooker'] == 0.7 and meta['target']['d'] == 2
    fuzz_cuda.check(meta, hist, lp_factor=1e4, rtol=1e-7)       # decision words are compared exactly inside check()


@pytest.mark.xfail(strict=True, reason='70 % snooker steps amplify the 1-ulp dot-product order difference ~4x per step '
                                       '(identical decisions; logp 165x the 1e-12 tolerance after 35 iterations)')
def test_snooker_drift_case_common_tolerance():
    import fuzz_cuda
    meta, hist = fuzz_cuda.random_case(DRIFT_CASE)
    fuzz_cuda.check(meta, hist)
