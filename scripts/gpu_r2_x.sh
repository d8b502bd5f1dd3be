#!/bin/bash
# round-2 call "x": ncu launch list of the final World bench command (per-launch durations, cold-cache / serialised: shares)
mkdir -p gpurun_out
BENCH="python bench.py --workload twitter-world --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
timeout 600 $BENCH > gpurun_out/x_bench_plain.json 2> gpurun_out/x_bench_plain.log && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/x_launches_world.csv $BENCH > gpurun_out/x_ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/x_bench_plain.log | cut -c1-200; wc -l gpurun_out/x_launches_world.csv
