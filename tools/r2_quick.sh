# quick GPU check: selected tests verbosely, then bench + phase stamps
T=${1:-r2q}
set -x
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_tc.py -m gpu -x -q -s -k "${2:-render or tc}" > gpurun_out/${T}_tests.log 2>&1; tail -25 gpurun_out/${T}_tests.log | cut -c1-400
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/'+'$T'+'_bench.json'))
    print(d['kernels_ms_per_step'], d['parity_on_sample'])
except Exception as e: print(e)
PY
timeout 300 python tools/phase_prof.py > gpurun_out/${T}_phase.log 2>&1; tail -25 gpurun_out/${T}_phase.log
