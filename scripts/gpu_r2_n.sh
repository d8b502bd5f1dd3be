#!/bin/bash
# round 2, call N (1 GPU): BASELINE config 5 sweep part 2; then the bench plain, its ncu launch list, and --set full
# captures of the headline SpMM kernel and the tcgen05 GEMM (CSV exports only: gpurun_out <= 64 MiB)
mkdir -p gpurun_out
for n in 5000000 10000000; do
  timeout 1500 python scripts/spmm_sweep.py --n $n --deg 5 10 20 50 --F 64 128 256 300 600 --graph chunglu community-sorted --panel -9 >> gpurun_out/n_config5_sweep.jsonl 2>> gpurun_out/n_config5_sweep.err
  echo "sweep n=$n rc=$?"
done
wc -l gpurun_out/n_config5_sweep.jsonl
BENCH="python bench.py --workload twitter-world --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
timeout 900 $BENCH > gpurun_out/n_bench_plain.json 2> gpurun_out/n_bench_plain.log && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/n_launches_world.csv $BENCH > gpurun_out/n_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:spmm_stream_kernel" -s 4 -c 2 -f -o /tmp/n_spmm $BENCH > gpurun_out/n_ncu_spmm.log 2>&1
echo "ncu spmm rc=$?"
ncu -i /tmp/n_spmm.ncu-rep --page raw --csv > gpurun_out/n_ncu_spmm_stream.raw.csv 2>/dev/null
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc_kernel" -s 2 -c 6 -f -o /tmp/n_gemm $BENCH > gpurun_out/n_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ncu -i /tmp/n_gemm.ncu-rep --page raw --csv > gpurun_out/n_ncu_gemm_tc.raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
