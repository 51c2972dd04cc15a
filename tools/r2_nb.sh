# quick GPU check of a kernel change: render / exchange parity tests, short bench with per-kernel times, phase stamps
T=${1:-r2n}
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_exchange.py -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log | cut -c1-300
timeout 200 python bench.py --steps 3 --warmup 2 --cpu-rays 0 --cpu-match-n3 0 2>gpurun_out/${T}_b1.err | python -c "
import json,sys
l=sys.stdin.readline()
try:
    d=json.loads(l); print('bench', d['value'], d['kernels_ms_per_step'])
except Exception as e: print('FAILED', l[:200])"; tail -2 gpurun_out/${T}_b1.err
timeout 200 python tools/phase_prof.py > gpurun_out/${T}_phase.log 2>&1; grep -E "^nb" gpurun_out/${T}_phase.log
