"""bench.py's output contract for the CPU (reference) arm, which needs no GPU: one JSON line with the keys the driver
reads; ranks other than 0 of a torchrun launch print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '1', '--steps', '6',
                           '--warmup', '3', '--cpu-budget', '5'], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_line():
    p = _run({})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == 'chain-steps/sec' and j['unit'] == 'chain-steps/s'
    for key in ('value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
                'config', 'cpu_baseline', 'e2e'):
        assert key in j, key
    assert j['value'] > 0 and j['higher_is_better'] is True and j['vs_baseline'] is None and j['dtype'] == 'f64'
    assert 'workload' in j['config'] and j['config']['ndim'] == 100 and j['config']['nchains'] == 1024
    assert j['steps'] == 6 and j['warmup'] == 3 and j['config']['iters_per_step'] == 1000      # the main arm's step definition
    assert 5 <= j['sampled_iterations_per_step'] <= 1000
    cb = j['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == j['value'] and cb['sample']
    assert j['e2e'] == dict(value=j['value'], unit=j['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_are_silent():
    p = _run(dict(RANK='1', WORLD_SIZE='2', LOCAL_RANK='1'))
    assert p.returncode == 0 and p.stdout.strip() == ''
