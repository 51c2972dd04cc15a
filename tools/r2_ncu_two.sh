T=${1:-r2p}
for K in neighbor2_kernel ray2_kernel; do
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -f -o gpurun_out/${T}_$K python bench.py --steps 1 --warmup 1 --cpu-rays 0 --cpu-match-n3 0 --rays 75776 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100
done
ls -la gpurun_out/${T}_*
