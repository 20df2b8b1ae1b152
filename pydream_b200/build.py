"""In-tree build of libdreamzs.so (hand-written sm_100a kernels + C ABI) with nvcc.

``python -m pydream_b200.build`` or ``pydream_b200.build.build()``.  Objects are compiled in
parallel (one translation unit per <G, R> kernel variant) and cached by mtime under
``pydream_b200/csrc/_obj``; the shared library lands at ``pydream_b200/libdreamzs.so`` so that
it travels with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
LIB = os.path.join(HERE, 'libdreamzs.so')

STEP_VARIANTS = [(4, 1), (8, 1), (16, 1), (32, 1), (32, 2), (32, 4), (32, 8)]
GAUSS_VARIANTS = [7, 8]
GWIN_VARIANTS = [7, 8]
MTP_VARIANTS = [(2, 1), (2, 2), (4, 1), (4, 2), (8, 1), (8, 2)]
PLAIN_UNITS = ['dreamzs_cabi.cu', 'dreamzs_adapt.cu', 'dreamzs_gr.cu', 'dreamzs_pt.cu', 'dreamzs_wwin_inst.cu']

# -fmad=false: parity-sensitive element-wise arithmetic must round like numpy;
# reductions use explicit fma().
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-fmad=false', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libdreamzs.so cannot be built')
    return nvcc


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'dreamzs.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(job):
    src, obj, defs, hdr_mtime, force = job
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime):
        return obj, None
    cmd = [_nvcc()] + NVCC_FLAGS + defs + ['-c', src, '-o', obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdr_mtime = _deps()
    jobs = []
    for g, r in STEP_VARIANTS:
        jobs.append((os.path.join(CSRC, 'dreamzs_step_inst.cu'), os.path.join(OBJ, 'step_%d_%d.o' % (g, r)),
                     ['-DDZ_G=%d' % g, '-DDZ_R=%d' % r], hdr_mtime, force))
    for tc in GAUSS_VARIANTS:
        jobs.append((os.path.join(CSRC, 'dreamzs_gauss_inst.cu'), os.path.join(OBJ, 'gauss_%d.o' % tc),
                     ['-DDZ_TC=%d' % tc], hdr_mtime, force))
    for tc in GWIN_VARIANTS:
        jobs.append((os.path.join(CSRC, 'dreamzs_gwin_inst.cu'), os.path.join(OBJ, 'gwin_%d.o' % tc),
                     ['-DDZ_TC=%d' % tc], hdr_mtime, force))
    for g, r in STEP_VARIANTS:
        jobs.append((os.path.join(CSRC, 'dreamzs_st2_inst.cu'), os.path.join(OBJ, 'st2_%d_%d.o' % (g, r)),
                     ['-DDZ_G=%d' % g, '-DDZ_R=%d' % r], hdr_mtime, force))
    for g, r in MTP_VARIANTS:
        jobs.append((os.path.join(CSRC, 'dreamzs_mtp_inst.cu'), os.path.join(OBJ, 'mtp_%d_%d.o' % (g, r)),
                     ['-DDZ_G=%d' % g, '-DDZ_R=%d' % r], hdr_mtime, force))
    for u in PLAIN_UNITS:
        jobs.append((os.path.join(CSRC, u), os.path.join(OBJ, u.replace('.cu', '.o')), [], hdr_mtime, force))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(_compile, jobs))
    objs = [o for o, _ in results]
    rebuilt = any(log is not None for _, log in results)
    if verbose:
        for o, log in results:
            if log:
                sys.stderr.write('== %s\n%s\n' % (os.path.basename(o), log))
    if rebuilt or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (p.stdout, p.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
