mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
echo "== two-level"; timeout 100 python tools/profile_step.py --iters 20010 --time 2>&1 | tail -1
echo "== one-level"; DREAMZS_WW_ONELEVEL=1 timeout 100 python tools/profile_step.py --iters 20010 --time 2>&1 | tail -1
done
echo "== two-level phases"; timeout 100 python tools/profile_step.py --iters 20010 --time --phases 2>&1 | tail -8 | head -2
echo "== c5"; timeout 100 python tools/profile_step.py --iters 1010 --time --chains 65536 --dim 50 --nseed 524288 2>&1 | tail -1
