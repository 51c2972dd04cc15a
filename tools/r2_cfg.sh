# configs[3] and the sweep on one GPU (short), multi-frame exchange test on however many GPUs are visible
T=${1:-r2y}
timeout 600 python bench.py --config 3 --steps 1 --warmup 1 --cpu-rays 0 > gpurun_out/${T}_cfg3.json 2> gpurun_out/${T}_cfg3.err; cut -c1-260 gpurun_out/${T}_cfg3.json; tail -2 gpurun_out/${T}_cfg3.err
timeout 900 python bench.py --config sweep --steps 1 --warmup 1 --sweep-max ${2:-20} > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err; python - <<PY
import json
for l in open('gpurun_out/${T}_sweep.jsonl'):
    d=json.loads(l); print(d['config']['sweep_point'], round(d['value']), round(d['ms_per_step'],1))
PY
tail -2 gpurun_out/${T}_sweep.err
