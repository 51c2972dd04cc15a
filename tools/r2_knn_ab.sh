set -x
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "knn" 2>&1 | tail -3
NLB_KNN_V1=1 timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "knn_ray" 2>&1 | tail -3
B="python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['value'], d['kernels_ms_per_step'])"; }
$B 2>/dev/null | k g8_seg16
NLB_KNN_SEG=8 $B 2>/dev/null | k g8_seg8
NLB_KNN_SEG=32 $B 2>/dev/null | k g8_seg32
NLB_KNN_SEG=64 $B 2>/dev/null | k g8_seg64
NLB_KNN_SEG=128 $B 2>/dev/null | k g8_seg128
NLB_KNN_V1=1 $B 2>/dev/null | k v1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
