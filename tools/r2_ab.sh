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
NLB_KNN_V1=1 $B 2>/dev/null | k v1_l8_f8
for v in l8_f4 l8_f2 l8_f3 l6_f4 l12_f4 l10_f4; do
NLB_LIB=$PWD/build/lib_$v.so NLB_KNN_V1=1 $B 2>/dev/null | k v1_$v
done
python bench.py --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('full', d['value'], d['kernels_ms_per_step'], d['parity_on_sample'])
except Exception as e: print('full FAILED', l[:200])"
timeout 300 python tools/phase_prof.py 2>&1 | sed -n 20,45p
