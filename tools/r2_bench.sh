T=${1:-r2t}
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').readline())
print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
print('kernels', d['kernels_ms_per_step'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('per_kernel','note')})
print('cpu', d['cpu_baseline'])
print('parity', d.get('parity_on_sample'))
print('setup', d['config']['per_frame_setup_ms'])
print('knn_cmp', d.get('knn_vs_reference_gpu'))
print('match', {k:v for k,v in d['match'].items() if k!='roofline'}); print(d['match']['roofline'])
print('pipeline', d.get('pipeline'))
PY
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null; cut -c1-700 gpurun_out/${T}_bench_ref.json
timeout 200 python tools/phase_prof.py > gpurun_out/${T}_phase.log 2>&1; grep -E "^ray|^nb2|^agg" gpurun_out/${T}_phase.log
