N=${1:-2}; T=${2:-r2z}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "rc=$?"; wc -c gpurun_out/${T}_bench_n$N.json; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${T}_bench_n$N.json').readline())
    print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'], d['config']['parallelism'][:80]); print(d['kernels_ms_per_step'])
except Exception as e: print('FAILED', e)
PY
tail -25 gpurun_out/${T}_bench_n$N.err | cut -c1-250
