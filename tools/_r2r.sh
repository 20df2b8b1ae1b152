mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multitry.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2r_pytest_mt.log
C3="--iters 410 --time --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5"
( echo "== two-stage mb4"; timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1
  for mb in 5 6; do echo "== minblocks $mb"; DREAMZS_LIB=$PWD/build/variants/libdreamzs_mtp_mb$mb.so timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1; done
) > gpurun_out/r2r_mtp.log 2>&1
cat gpurun_out/r2r_pytest_mt.log gpurun_out/r2r_mtp.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mt -s 6 -c 2 -o gpurun_out/r2r_mt2 -f python tools/profile_step.py --iters 60 --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5 > gpurun_out/r2r_ncu.log 2>&1; tail -2 gpurun_out/r2r_ncu.log
