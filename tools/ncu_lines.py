"""Per-source-line stall samples of the first launch in an ncu report (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [min_pct]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-count', '1', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
lines = []   # (file, line, src, samples, inst)
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif len(r) > 6 and r[0] == 'Line No':
        hdr = r
        isamp, iinst = hdr.index('# Samples'), hdr.index('Instructions Executed')
    elif hdr and len(r) > isamp and r[0] not in ('', 'Line No'):
        try:
            lines.append((cur_file, int(r[0]), r[1].strip()[:100], int(r[isamp]), int(r[iinst])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines)
toti = sum(l[4] for l in lines)
print('total samples %d, warp-instructions %d' % (tot, toti))
for f, ln, src, s, ni in lines:
    if s >= tot * minpct / 100:
        print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100 * s / tot, 100 * ni / toti, f, ln, src))

if len(sys.argv) > 3:   # bucket boundaries "file:line,line,line..."
    fname, bounds = sys.argv[3].split(':')
    bounds = [int(b) for b in bounds.split(',')]
    buckets = {}
    for f, ln, src, s, ni in lines:
        if f == fname:
            k = sum(1 for b in bounds if ln >= b)
            key = '%s:%d-' % (fname, bounds[k - 1] if k else 0)
        else:
            key = f
        a = buckets.setdefault(key, [0, 0])
        a[0] += s
        a[1] += ni
    print('--- buckets')
    for k, (s, ni) in sorted(buckets.items()):
        print('%5.1f%% smp %5.1f%% inst  %s' % (100 * s / tot, 100 * ni / toti, k))
