#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check exit $?"; grep -v "^tf32 " gpurun_out/tc_check.log | tail -45
for wl in twitter-us twitter-world; do
timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log
echo "bench $wl exit $?"; tail -2 gpurun_out/bench_$wl.log; cat gpurun_out/bench_$wl.json
done
