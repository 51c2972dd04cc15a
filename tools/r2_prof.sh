T=${1:-r2f}
timeout 300 python tools/phase_prof.py > gpurun_out/${T}_phase.log 2>&1; cat gpurun_out/${T}_phase.log | head -60
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_query_rays|aggregate_kernel|neighbor2_kernel|row_gemm128|ray2_kernel" -c 96 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"knn_query_rays|aggregate_kernel|neighbor2_kernel|row_gemm128|ray2_kernel" -s 6 -c 6 -f -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/${T}_*
