# ncu --set full of one kernel (regex $2), report named $1
T=${1:-r2k}; K=${2:-aggregate_kernel}
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -f -o gpurun_out/${T}_full python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/${T}_full.ncu-rep
