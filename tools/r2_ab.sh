T=${1:-r2i}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -30 gpurun_out/${T}_tests.log | cut -c1-300
B="python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0"
k() { python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('$1', d['value'], d['kernels_ms_per_step'], d['config']['per_frame_setup_ms'])
except Exception as e: print('$1', 'FAILED', l[:200])"; }
$B 2>gpurun_out/${T}_b1.err | k default; tail -3 gpurun_out/${T}_b1.err
NLB_AGG_V1=1 $B 2>/dev/null | k agg_v1
NLB_KNN_V1=1 NLB_LIB=$PWD/build/lib_l8_f4.so $B 2>/dev/null | k knn_l8f4
python bench.py --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('full', d['value'], d['kernels_ms_per_step'], d['parity_on_sample'])
except Exception as e: print('full FAILED', l[:200])"
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_query|visibility_kernel|aggregate_kernel|neighbor2_kernel|row_gemm128|ray2_kernel" -c 70 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-100
