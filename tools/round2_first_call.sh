#!/bin/bash
# The first GPU call of the next round, prepared at the end of round 1 (GPU budget spent): A/B of the window-kernel
# variants that compile but are not measured yet, phase stamps of the staggered build, sanitizer on the tempering kernels.
#   here (no GPU):   bash tools/round2_first_call.sh build
#   on the B200:     gpurun --timeout 600 -- 'bash tools/round2_first_call.sh run'
set -u
cd "$(dirname "$0")/.."
V=build/variants
case "${1:-}" in
build)
  python tools/build_variant.py stagger1 -DDZ_GW_STAGGER=1
  python tools/build_variant.py stagger2 -DDZ_GW_STAGGER=2
  python tools/build_variant.py stagger3 -DDZ_GW_STAGGER=3
  python tools/build_variant.py fastnormal -DDZ_FAST_NORMAL=1
  python tools/build_variant.py stagger2fn -DDZ_GW_STAGGER=2 -DDZ_FAST_NORMAL=1
  ;;
run)
  mkdir -p gpurun_out
  # 1. A/B, interleaved child processes, parity of decisions / logp against the shipped build
  timeout 400 python tools/ab_libs.py --iters 3000 --reps 2 base=pydream_b200/libdreamzs.so stagger1=$V/libdreamzs_stagger1.so \
      stagger2=$V/libdreamzs_stagger2.so stagger3=$V/libdreamzs_stagger3.so fastnormal=$V/libdreamzs_fastnormal.so \
      stagger2fn=$V/libdreamzs_stagger2fn.so 2>&1 | tee gpurun_out/r2a_ab.log
  # 2. phase stamps of both warp groups, shipped vs staggered
  DREAMZS_LIB=$PWD/pydream_b200/libdreamzs.so timeout 60 python tools/profile_step.py --iters 41 --phases 2>&1 | tail -3 | tee gpurun_out/r2a_phases_base.log
  DREAMZS_LIB=$PWD/$V/libdreamzs_stagger2.so timeout 60 python tools/profile_step.py --iters 41 --phases 2>&1 | tail -4 | tee gpurun_out/r2a_phases_stagger2.log
  # 3. window-kernel parity tests on the best candidate (edit the name after step 1)
  DREAMZS_LIB=$PWD/$V/libdreamzs_stagger2.so timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2a_parity_stagger2.log
  # 4. memcheck on the tempering kernels (small case)
  timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tempering.py -q -m gpu -k "golden" 2>&1 | tail -8 | tee gpurun_out/r2a_sanitizer_pt.log
  ;;
*)
  echo "usage: $0 build|run"; exit 2 ;;
esac
