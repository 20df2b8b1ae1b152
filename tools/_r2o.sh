mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2o_pytest_gpu.log
C3="--iters 410 --time --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5"
( echo "== default (minblocks 4)"; timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1
  echo "== generic"; timeout 100 python tools/profile_step.py $C3 --generic 2>&1 | tail -1
  for mb in 3 5 6; do echo "== minblocks $mb"; DREAMZS_LIB=$PWD/build/variants/libdreamzs_mtp_mb$mb.so timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1; done
  echo "== d=20 mt3 gaussian 4096"; timeout 100 python tools/profile_step.py --iters 410 --time --chains 4096 --dim 20 --nseed 1048576 --multitry 3 2>&1 | tail -1
  echo "== d=20 mt3 gaussian 4096 generic"; timeout 100 python tools/profile_step.py --iters 410 --time --chains 4096 --dim 20 --nseed 1048576 --multitry 3 --generic 2>&1 | tail -1
) > gpurun_out/r2o_mtp.log 2>&1
cat gpurun_out/r2o_pytest_gpu.log gpurun_out/r2o_mtp.log
