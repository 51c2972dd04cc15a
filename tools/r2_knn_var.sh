B="python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$1', d['value'], d['kernels_ms_per_step'], d['config']['per_frame_setup_ms'])
except Exception as e: print('$1', 'FAILED', l[:200])"; }
NLB_KNN_V1=1 $B 2>/dev/null | k v1_l8_f8
for v in l4_f8 l8_f4 l4_f4 l16_f8; do
NLB_LIB=$PWD/build/lib_$v.so NLB_KNN_V1=1 $B 2>/dev/null | k v1_$v
done
NLB_LIB=$PWD/build/lib_l4_f8.so NLB_KNN_V1=1 timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "knn" 2>&1 | tail -2
