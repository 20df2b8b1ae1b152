mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multitry.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2s_pytest_mt.log
C3="--iters 410 --time --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5"
( echo "== two-stage"; timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1
  timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1
) > gpurun_out/r2s_mtp.log 2>&1
cat gpurun_out/r2s_pytest_mt.log gpurun_out/r2s_mtp.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2s_launches.csv python tools/profile_step.py --iters 60 --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5 > /dev/null 2>&1
