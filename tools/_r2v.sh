for i in 1 2; do
echo "== dmma sums"; timeout 100 python tools/profile_step.py --iters 10010 --time 2>&1 | tail -1
echo "== shuffle sums"; DREAMZS_LIB=$PWD/build/variants/libdreamzs_shflsum.so timeout 100 python tools/profile_step.py --iters 10010 --time 2>&1 | tail -1
done
echo "== shuffle sums phases"; DREAMZS_LIB=$PWD/build/variants/libdreamzs_shflsum.so timeout 100 python tools/profile_step.py --iters 10010 --time --phases 2>&1 | tail -8 | head -4
DREAMZS_LIB=$PWD/build/variants/libdreamzs_shflsum.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
