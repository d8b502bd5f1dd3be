#!/bin/bash
# ncu launch list of one bench command + full captures of the headline kernels (1 GPU).
WL=${1:-twitter-world}; TAG=${2:-r01}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$WL.json 2> gpurun_out/plain_$WL.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_$WL.csv $CMD > gpurun_out/ncu_list_$WL.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/plain2_$WL.json 2> gpurun_out/plain2_$WL.log &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_vec_kernel_nb|gemm_tc_kernel" -s 30 -c 6 -f -o gpurun_out/${TAG}_kernels_$WL $CMD > gpurun_out/ncu_full_$WL.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/ | grep $TAG
