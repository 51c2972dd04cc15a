T=${1:-r2i}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -6 gpurun_out/${T}_tests.log | cut -c1-300
B="timeout 150 python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$1', d['value'], d['kernels_ms_per_step'], d['config']['per_frame_setup_ms'])
except Exception as e: print('$1', 'FAILED', l[:200])"; }
$B 2>gpurun_out/${T}_b1.err | k default; tail -3 gpurun_out/${T}_b1.err
NLB_AGG_FC_V1=1 $B 2>/dev/null | k fc_inside
timeout 300 python bench.py --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('full', d['value'], d['kernels_ms_per_step'], d['parity_on_sample'])
except Exception as e: print('full FAILED', l[:200])"
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_query|visibility_kernel|aggregate_kernel|neighbor2_kernel|row_gemm128|ray2_kernel|fc_tail" -c 70 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-100
python - <<PY
import csv,collections
rows=list(csv.reader(open('gpurun_out/${T}_launches.csv')))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d.get('Metric Name')=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
            if u=='ns': v/=1e6
            elif u=='us': v/=1e3
            agg[d['Kernel Name'][:50]].append(v)
for k,v in agg.items(): print(f"{k:50s} n={len(v):3d} mean={sum(v)/len(v):7.3f} ms  x8.1 = {sum(v)/len(v)*8.108:7.1f} ms/frame")
PY
