N=${1:-8}; T=${2:-r2z}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "rc=$?"
timeout 400 $TR bench.py --gpus $N --config 3 --steps 3 --warmup 2 > gpurun_out/${T}_cfg3_n$N.json 2> gpurun_out/${T}_cfg3_n$N.err; echo "rc=$?"
timeout 500 $TR bench.py --gpus $N --config sweep --steps 1 --warmup 1 > gpurun_out/${T}_sweep_n$N.jsonl 2> gpurun_out/${T}_sweep_n$N.err; echo "rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${T}_bench_n$N.json','gpurun_out/${T}_cfg3_n$N.json','gpurun_out/${T}_sweep_n$N.jsonl'):
    for l in open(f):
        l=l.strip()
        if l.startswith('{'):
            d=json.loads(l); print(f.split('/')[-1], d['config'].get('sweep_point'), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1))
PY
tail -3 gpurun_out/${T}_sweep_n$N.err | cut -c1-200
