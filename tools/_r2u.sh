C3="--iters 410 --time --chains 4096 --dim 10 --nseed 2097152 --target mixture --multitry 5"
for mb in 5 6 7; do echo "== minblocks $mb"; DREAMZS_LIB=$PWD/build/variants/libdreamzs_mtp_mb$mb.so timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1; done
echo "== base"; timeout 100 python tools/profile_step.py $C3 2>&1 | tail -1
