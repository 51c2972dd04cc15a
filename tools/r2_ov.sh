B="timeout 150 python bench.py --steps 3 --warmup 2 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$1', d['value'], d['ms_per_step'], sum(d['kernels_ms_per_step'].values()), d['e2e']['value'])
except Exception as e: print('$1', 'FAILED', l[:200])"; }
$B 2>/dev/null | k no_overlap
NLB_KNN_OVERLAP=1 $B 2>/dev/null | k overlap
$B --chunk 18944 2>/dev/null | k chunk18944
$B --chunk 75776 2>/dev/null | k chunk75776
