C5="--iters 1010 --time --chains 65536 --dim 50 --nseed 524288"
for tc in 8 12 16 24 32; do echo "== TC=$tc"; DREAMZS_WW_TC=$tc timeout 100 python tools/profile_step.py $C5 2>&1 | tail -1; done
echo "== TC=16 NB=5"; DREAMZS_WW_TC=16 DREAMZS_WW_NB=5 timeout 100 python tools/profile_step.py $C5 2>&1 | tail -1
echo "== default phases"; timeout 100 python tools/profile_step.py $C5 --phases 2>&1 | tail -8 | cut -c1-400
echo "== nopersist"; timeout 100 python tools/profile_step.py $C5 --nopersist 2>&1 | tail -1
