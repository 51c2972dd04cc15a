set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1i_tests.log 2>&1; tail -2 gpurun_out/r1i_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1; tail -2 gpurun_out/r1i_smoke.log
python bench.py > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err; cut -c1-160 gpurun_out/r1i_bench.json; tail -2 gpurun_out/r1i_bench.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1i_bench_ref.json 2>/dev/null; cut -c1-160 gpurun_out/r1i_bench_ref.json
python tools/phase_prof.py > gpurun_out/r1i_phase.log 2>&1
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_query_rays|aggregate_kernel|neighbor_kernel|ray_kernel" -c 64 --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log | cut -c1-200
timeout 280 ncu --set full --clock-control none --import-source on -k regex:"knn_query_rays|aggregate_kernel|neighbor_kernel|ray_kernel" -s 4 -c 4 -f -o gpurun_out/r1i_full python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/r1i_*
