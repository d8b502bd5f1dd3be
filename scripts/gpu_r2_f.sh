#!/bin/bash
# round 2, call F: chunked accumulation chains in the tcgen05 GEMM -- tests, rates, World parity/bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mlpconv.py tests/test_gpu_layers.py -m gpu -x -q --timeout 300 > gpurun_out/f_pytest_gemm.log 2>&1
echo "pytest gemm rc=$?"; tail -5 gpurun_out/f_pytest_gemm.log
timeout 600 python scripts/tc_check.py > gpurun_out/f_tc_check.log 2>&1; echo "tc_check rc=$?"; grep "tf32x3" gpurun_out/f_tc_check.log | tail -6
timeout 1200 python bench.py --workload twitter-world --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_world.json 2> gpurun_out/f_bench_world.log
echo "bench rc=$?"; grep "parity" gpurun_out/f_bench_world.log | awk '{ for(i=1;i<=NF;i++) if ($i=="scaled") v=$(i+1); print v, $0 }' | sort -n -r | head -5 | cut -c1-200; tail -2 gpurun_out/f_bench_world.log | cut -c1-300
