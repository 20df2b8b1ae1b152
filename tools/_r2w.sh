mkdir -p gpurun_out
N=${1:-2}
timeout 200 python -m pytest tests/test_multigpu.py -m gpu -q -x -k nvlink 2>&1 | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err
tail -3 gpurun_out/r2w_bench_n$N.err; python - <<PY
import json
for line in open('gpurun_out/r2w_bench_n$N.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['check'].get('sharded_equals_single'))
PY
