set -x
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "knn" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['value'], d['kernels_ms_per_step'])"; }
$B 2>/dev/null | k default
NLB_NB_V1=1 $B 2>/dev/null | k nb_v1
NLB_KNN_V1=1 $B 2>/dev/null | k knn_v1
NLB_KNN_SEG=8 $B 2>/dev/null | k g8_seg8
NLB_KNN_SEG=32 $B 2>/dev/null | k g8_seg32
NLB_KNN_SEG=64 $B 2>/dev/null | k g8_seg64
python bench.py --steps 2 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('full', d['value'], d['kernels_ms_per_step'], d['parity_on_sample'])"
