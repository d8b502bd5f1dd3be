#!/bin/bash
# round 2, call E: GEMM accuracy (separate TMEM accumulator for the cross terms) + epoch breakdown + GEMM tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mlpconv.py tests/test_gpu_layers.py -m gpu -x -q > gpurun_out/e_pytest_gemm.log 2>&1
echo "pytest gemm rc=$?"; tail -5 gpurun_out/e_pytest_gemm.log
timeout 600 python scripts/tc_check.py > gpurun_out/e_tc_check.log 2>&1; echo "tc_check rc=$?"; tail -12 gpurun_out/e_tc_check.log
timeout 1200 python bench.py --workload twitter-world --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/e_bench_world.json 2> gpurun_out/e_bench_world.log
echo "bench rc=$?"; grep -A16 "op breakdown" gpurun_out/e_bench_world.log; grep "parity" gpurun_out/e_bench_world.log | sort -k4 -n -r | head -8; tail -2 gpurun_out/e_bench_world.log
