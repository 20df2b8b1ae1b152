mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2v_pytest_gpu.log
( timeout 100 python tools/profile_step.py --iters 20010 --time --phases 2>&1 | tail -8
  timeout 100 python tools/profile_step.py --iters 20010 --time 2>&1 | tail -1
  timeout 100 python tools/profile_step.py --iters 1010 --time --chains 65536 --dim 50 --nseed 524288 2>&1 | tail -1 ) > gpurun_out/r2v_time.log 2>&1
cat gpurun_out/r2v_pytest_gpu.log gpurun_out/r2v_time.log
