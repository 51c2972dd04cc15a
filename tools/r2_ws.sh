B="timeout 150 python bench.py --steps 2 --warmup 2 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$1', d['value'], d['kernels_ms_per_step']['ray'])
except Exception as e: print('$1', 'FAILED', l[:200])"; }
$B 2>/dev/null | k ws1
NLB_R2_WSPLIT=2 $B 2>/dev/null | k ws2
NLB_R2_WSPLIT=4 $B 2>/dev/null | k ws4
NLB_R2_WSPLIT=8 $B 2>/dev/null | k ws8
