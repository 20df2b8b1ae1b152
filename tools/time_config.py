"""Device-resident timing of one BASELINE config with bench.py's own timed regions (pre-allocated trace, repeated regions):
    python tools/time_config.py c4 [--generic]        # prints value, ms per step, launches"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == '__main__':
    name = sys.argv[1]
    if '--generic' in sys.argv:
        bench.WORKLOADS[name]['opts'] = dict(bench.WORKLOADS[name]['opts'], generic_kernel=True)
    out = bench.bench_other_config(name, 5, 3)
    print('%s: %.1f M chain-steps/s, %.3f ms per step of %d iterations, %d launches, roofline %.3f'
          % (name, out['value'] / 1e6, out['ms_per_step'], out['iters_per_step'], out['gpu_launches'], out['roofline']['frac']))
