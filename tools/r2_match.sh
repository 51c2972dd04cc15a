T=${1:-r2s}
timeout 600 python -m pytest tests/test_gpu_matcher.py -m gpu -x -q -s > gpurun_out/${T}_tests.log 2>&1; tail -8 gpurun_out/${T}_tests.log | cut -c1-300
timeout 400 python bench.py --steps 2 --warmup 2 2>gpurun_out/${T}_b.err | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('full', d['value'], d['kernels_ms_per_step'], d['parity_on_sample']); print(d['match'])
except Exception as e: print('full FAILED', l[:300])"
tail -3 gpurun_out/${T}_b.err
NLB_S2D_V1=1 timeout 400 python bench.py --steps 1 --warmup 1 --cpu-rays 256 2>/dev/null | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('s2d_v1', d['match'])
except Exception as e: print('FAILED', l[:300])"
