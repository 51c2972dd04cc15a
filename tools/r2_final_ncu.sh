# one gpurun call: GPU tests, default bench line, ncu launch list of one step, ncu --set full of every render kernel
# (the report stays on the box - it exceeds what gpurun brings back - and its raw page is exported as CSV)
T=${1:-r2x}
R="knn_query_rays|visibility_kernel|aggregate_kernel|fc_tail_kernel|neighbor2_kernel|row_gemm128|ray2_kernel"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$R" -c 96 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log | cut -c1-200
timeout 700 ncu --set full --clock-control none -k regex:"$R" -c 18 -f -o /tmp/${T}_full python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
timeout 120 ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw.csv 2> gpurun_out/ncu_export.err
timeout 300 python tools/phase_prof.py > gpurun_out/${T}_phase.log 2>&1; tail -5 gpurun_out/${T}_phase.log
ls -la gpurun_out/${T}_* /tmp/${T}_full.ncu-rep
