# one gpurun call: tests, bench, phase stamps (names the outputs after $1)
T=${1:-r2a}
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -5 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 300 python tools/phase_prof.py > gpurun_out/${T}_phase.log 2>&1; tail -40 gpurun_out/${T}_phase.log
