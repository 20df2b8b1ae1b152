N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
if [ "$N" = 2 ]; then timeout 250 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -2; fi
for i in 1 2 3 4; do echo "== peers run $i (no phases)"; timeout 60 $TR tools/profile_multi.py 2>&1 | grep "rank 0"; done
echo "== peers phases"; timeout 60 $TR tools/profile_multi.py --phases 2>&1 | grep -A1 "rank 0"
