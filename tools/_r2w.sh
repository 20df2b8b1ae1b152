mkdir -p gpurun_out
N=${1:-2}; WL=${2:-c2}; EI=${3:-2000}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --workload $WL --e2e-iters $EI > gpurun_out/r2w_bench_${WL}_n$N.json 2> gpurun_out/r2w_bench_${WL}_n$N.err
tail -3 gpurun_out/r2w_bench_${WL}_n$N.err; python - <<PY
import json
for line in open('gpurun_out/r2w_bench_${WL}_n$N.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['n_gpus'], d['config']['workload'][:40], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['check'].get('sharded_equals_single'))
PY
