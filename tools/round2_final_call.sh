#!/bin/bash
# Round-2 measurements on one B200 (gpurun --timeout 1800 -- 'bash tools/round2_final_call.sh'): GPU test suite, bench lines
# of both arms, ncu launch lists + full-set captures of the dominant kernels (C2: whitened window kernel, C3: two-stage
# multi-try kernels), sanitizer runs of the window / multi-try / tempering kernels on small cases.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2z_pytest_gpu.log
timeout 600 python bench.py > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2z_bench_reference_n1.json 2> $O/r2z_bench_reference_n1.err
C2="--iters 2011"
C3="--iters 60 --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5"
C5="--iters 111 --chains 65536 --dim 50 --nseed 524288"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2z_launches_c2.csv python tools/profile_step.py $C2 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2z_launches_c3.csv python tools/profile_step.py $C3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wwin_kernel -s 1 -c 1 -o $O/r2z_wwin_c2 -f python tools/profile_step.py $C2 > $O/r2z_ncu_c2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mt -s 6 -c 3 -o $O/r2z_mt_c3 -f python tools/profile_step.py $C3 > $O/r2z_ncu_c3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wwin_kernel -s 1 -c 1 -o $O/r2z_wwin_c5 -f python tools/profile_step.py $C5 > $O/r2z_ncu_c5.log 2>&1
( timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "whitened or c2_gauss100" 2>&1 | tail -6
  timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_multitry.py -q -m gpu -k "mix10_mt5 or const4_mt3_regen or sum6_mt5_bounds" 2>&1 | tail -6
  timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_tempering.py -q -m gpu -k "golden" 2>&1 | tail -6
) > $O/r2z_sanitizer_memcheck.log 2>&1
( timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_multitry.py -q -m gpu -k "mix10_mt5-two_stage or gauss20_mt3-two_stage" 2>&1 | tail -6
  timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "whitened and d24" 2>&1 | tail -6
) > $O/r2z_sanitizer_racecheck.log 2>&1
tail -3 $O/r2z_pytest_gpu.log; tail -2 $O/r2z_bench_n1.err; tail -4 $O/r2z_sanitizer_memcheck.log; tail -4 $O/r2z_sanitizer_racecheck.log
