"""Differential fuzzing of the CUDA path (through the C ABI, pydream_b200.engine) against the C oracle over the random
option combinations of tools/fuzz_oracle.py (which pins the oracle itself to the unmodified reference): adaptation,
gamma levels, every closed-form prior, no hard boundaries, multi-try 3-5, DE pairs 1-3, parallel tempering.  Runs on a
B200 (no reference needed).      python tools/fuzz_cuda.py [ncases] [first_seed]"""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))

from fuzz_oracle import random_case                                          # noqa: E402
from golden_util import make_target, prior_arrays, sampler_kwargs, logp_tol  # noqa: E402
from oracle import c_oracle                                                  # noqa: E402


def check(meta, hist, lp_factor=10, rtol=1e-9):
    from pydream_b200.engine import DreamEngine
    d, N, T = meta['target']['d'], meta['N'], meta['T']
    starts = hist[:N].copy()
    tgt = make_target(meta['target'])
    pk, pa, pb = prior_arrays(meta['prior'], d)
    kw = sampler_kwargs(meta)
    orc = c_oracle.OracleSampler(d, N, hist, starts, tgt.kind, tgt.table(), seed=meta['seed'], prior_kind=pk, prior_a=pa,
                                 prior_b=pb, **kw)
    eng = DreamEngine(d, N, hist, starts, tgt, pk, pa, pb, seed=meta['seed'], **kw)
    if meta['tempering']:
        ref = orc.run_pt(T)
        trace, logp, dec, swaps = eng.run_tempered(T)
        assert np.array_equal(swaps[:, :3].cpu().numpy().astype(np.int64), ref['swaps']), 'swaps'
        got_dec, got_lp = dec.cpu().numpy().astype(np.uint32), logp.cpu().numpy()
        got_sp = trace[:, :, :d].cpu().numpy()
        ref_dec, ref_lp, ref_sp = ref['decisions'], ref['log_ps'], ref['sampled_params']
    else:
        ref = orc.run(T)
        trace, logp, dec = eng.run(T)
        got_dec = dec.t().contiguous().cpu().numpy().astype(np.uint32)
        got_lp = logp.t().contiguous().cpu().numpy()
        got_sp = trace[:, :, :d].permute(1, 0, 2).contiguous().cpu().numpy()
        ref_dec, ref_lp, ref_sp = ref['decisions'], ref['logp'], ref['states']
    assert np.array_equal(got_dec, ref_dec), 'decisions differ at %s' % (np.argwhere(got_dec != ref_dec)[:3].tolist(),)
    err = np.abs(got_lp - ref_lp) / logp_tol(ref_lp)
    assert np.all(err <= lp_factor), 'logp: %.2f x tolerance' % err.max()
    np.testing.assert_allclose(got_sp, ref_sp, rtol=rtol, atol=rtol / 10)
    np.testing.assert_allclose(eng.history_flat(), orc.history_flat, rtol=rtol, atol=rtol / 10)
    np.testing.assert_allclose(eng.cr_probs.cpu().numpy(), orc.cr_probs, rtol=rtol)
    np.testing.assert_allclose(eng.gamma_probs.cpu().numpy(), orc.gamma_probs, rtol=rtol)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for i in range(first, first + n):
        meta, hist = random_case(i)
        try:
            check(meta, hist)
        except Exception:      # noqa: BLE001
            bad += 1
            msg = traceback.format_exc().strip().splitlines()
            print('case %d FAILED: %s\n   %s' % (i, {k: meta[k] for k in ('target', 'N', 'T', 'kw', 'tempering')}, msg[-1][:400]), flush=True)
    print('%d cases, %d failed' % (n, bad))


if __name__ == '__main__':
    main()
