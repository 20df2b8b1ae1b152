"""Build an A/B variant of libdreamzs.so: the window-kernel translation units recompiled with extra -D flags, linked with
the objects of the default build (pydream_b200/csrc/_obj).  The result goes to build/variants/ (git-ignored, travels
to the GPU box) and is selected with DREAMZS_LIB / tools/ab_libs.py.
    python tools/build_variant.py stagger2 -DDZ_GW_STAGGER=2
    python tools/build_variant.py groups4 -DDZ_GW_GROUPS=4
    python tools/ab_libs.py base=pydream_b200/libdreamzs.so stagger2=build/variants/libdreamzs_stagger2.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pydream_b200 import build as B     # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    B.build()
    out = os.path.join(ROOT, 'build', 'variants')
    os.makedirs(out, exist_ok=True)
    if any(x.startswith('-DDZ_ST2') for x in defs):     # variants of the two-stage single-try kernels
        objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith('.o') and not f.startswith('st2_')]
        mine = []
        for g, r in B.STEP_VARIANTS:
            o = os.path.join(out, '%s_st2_%d_%d.o' % (name, g, r))
            cmd = [B._nvcc()] + B.NVCC_FLAGS + ['-DDZ_G=%d' % g, '-DDZ_R=%d' % r] + defs + ['-c', os.path.join(B.CSRC, 'dreamzs_st2_inst.cu'), '-o', o]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise SystemExit(p.stdout + p.stderr)
            mine.append(o)
        lib = os.path.join(out, 'libdreamzs_%s.so' % name)
        subprocess.check_call([B._nvcc(), '-shared', '-o', lib] + objs + mine + ['-gencode', 'arch=compute_100a,code=sm_100a'])
        for o in mine:
            os.remove(o)
        print(lib)
        return
    if any(x.startswith('-DDZ_MTP') for x in defs):     # variants of the point-parallel multi-try kernel
        objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith('.o') and not f.startswith('mtp_')]
        mine = []
        for g, r in B.MTP_VARIANTS:
            o = os.path.join(out, '%s_mtp_%d_%d.o' % (name, g, r))
            cmd = [B._nvcc()] + B.NVCC_FLAGS + ['-DDZ_G=%d' % g, '-DDZ_R=%d' % r] + defs + ['-c', os.path.join(B.CSRC, 'dreamzs_mtp_inst.cu'), '-o', o]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise SystemExit(p.stdout + p.stderr)
            for line in p.stderr.splitlines():
                if 'spill' in line or 'registers' in line:
                    print(line.strip()[:160])
            mine.append(o)
        lib = os.path.join(out, 'libdreamzs_%s.so' % name)
        subprocess.check_call([B._nvcc(), '-shared', '-o', lib] + objs + mine + ['-gencode', 'arch=compute_100a,code=sm_100a'])
        for o in mine:
            os.remove(o)
        print(lib)
        return
    if any(x.startswith('-DDZ_WW') for x in defs):      # variants of the whitened window kernel
        objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith('.o') and f != 'dreamzs_wwin_inst.o']
        o = os.path.join(out, '%s_wwin.o' % name)
        cmd = [B._nvcc()] + B.NVCC_FLAGS + defs + ['-c', os.path.join(B.CSRC, 'dreamzs_wwin_inst.cu'), '-o', o]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise SystemExit(p.stdout + p.stderr)
        for line in p.stderr.splitlines():
            if 'spill' in line or 'registers' in line:
                print(line.strip()[:160])
        lib = os.path.join(out, 'libdreamzs_%s.so' % name)
        subprocess.check_call([B._nvcc(), '-shared', '-o', lib] + objs + [o, '-gencode', 'arch=compute_100a,code=sm_100a'])
        os.remove(o)
        print(lib)
        return
    objs = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith('.o') and not f.startswith('gwin_')]
    for tc in B.GWIN_VARIANTS:
        o = os.path.join(out, '%s_gwin_%d.o' % (name, tc))
        cmd = [B._nvcc()] + B.NVCC_FLAGS + ['-DDZ_TC=%d' % tc] + defs + ['-c', os.path.join(B.CSRC, 'dreamzs_gwin_inst.cu'), '-o', o]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise SystemExit(p.stdout + p.stderr)
        for line in p.stderr.splitlines():
            if 'spill' in line or 'registers' in line:
                print(line.strip()[:160])
        objs.append(o)
    lib = os.path.join(out, 'libdreamzs_%s.so' % name)
    subprocess.check_call([B._nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    for tc in B.GWIN_VARIANTS:
        os.remove(os.path.join(out, '%s_gwin_%d.o' % (name, tc)))
    print(lib)


if __name__ == '__main__':
    main()
