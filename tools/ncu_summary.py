"""Summarise ncu outputs brought back from the GPU box (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv        # per-kernel share of the step
    python tools/ncu_summary.py report   gpurun_out/x.ncu-rep [regex]     # key metrics + SASS instruction mix
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum']


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        if row['Metric Unit'] in ('us', 'usecond'):
            v *= 1e3
        elif row['Metric Unit'] in ('ms', 'msecond'):
            v *= 1e6
        k = re.sub(r'\(.*', '', row['Kernel Name'])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(t for _, t in agg.values())
    print('%-70s %8s %12s %8s' % ('kernel', 'launches', 'avg ns', 'share'))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-70s %8d %12.1f %7.1f%%' % (k[:70], n, t / n, 100 * t / tot))


def report(path, pattern=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    name_i = h.index('Kernel Name')
    for r in rows[2:]:
        if pattern and not re.search(pattern, r[name_i]):
            continue
        print('== %s' % r[name_i][:100])
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print('  %-82s %16s %s' % (k, r[i], units[i]))
        stall = [(float(r[i]), n) for i, n in enumerate(h) if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio')]
        print('  top stalls (warps per issue-active): ' + ', '.join('%s %.2f' % (n.split('stalled_')[1].split('_per_')[0], v) for v, n in sorted(stall, reverse=True)[:6]))
        if pattern is not None:
            break
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    h = rows[hi]
    ia, isrc = h.index('Instructions Executed'), h.index('Source')
    ops = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= ia:
            continue
        t = r[isrc].split()
        if not t:
            continue
        o = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
        try:
            ops[o.split('.')[0]] += int(r[ia])
        except ValueError:      # header row of a further kernel
            break
    tot = sum(ops.values())
    print('  SASS warp-instructions executed (first launch): %d' % tot)
    print('  ' + ', '.join('%s %.1f%%' % (o, 100 * n / tot) for o, n in ops.most_common(14)))
    f64 = sum(n for o, n in ops.items() if o in ('DFMA', 'DADD', 'DMUL', 'DSETP', 'DMMA'))
    print('  fp64-pipe warp-instructions: %d (%.1f%%)' % (f64, 100 * f64 / tot))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
