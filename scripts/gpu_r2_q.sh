#!/bin/bash
# round 2, call Q (1 GPU): what the driver runs at round end -- smoke, the default bench line (with cpu_baseline and
# parity), the reference arm at the same configuration -- plus GEMM tests / rates of the final kernels
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/q_smoke.log
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mlp.py -m gpu -x -q --timeout 300 > gpurun_out/q_pytest_gemm.log 2>&1; echo "pytest gemm+mlp rc=$?"; tail -3 gpurun_out/q_pytest_gemm.log
timeout 600 python scripts/tc_check.py > gpurun_out/q_tc_check.log 2>&1; echo "tc_check rc=$?"; grep "bench tf32x3" gpurun_out/q_tc_check.log
( time timeout 1500 python bench.py --breakdown > gpurun_out/q_bench_default.json 2> gpurun_out/q_bench_default.log ) 2> gpurun_out/q_bench_default.time
echo "bench default rc=$?"; cat gpurun_out/q_bench_default.time | tail -3; grep -A12 "op breakdown" gpurun_out/q_bench_default.log | cut -c1-110
( time timeout 1700 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/q_bench_reference.json 2> gpurun_out/q_bench_reference.log ) 2> gpurun_out/q_bench_reference.time
echo "bench reference rc=$?"; tail -3 gpurun_out/q_bench_reference.time; tail -4 gpurun_out/q_bench_reference.log | cut -c1-200; cut -c1-400 gpurun_out/q_bench_reference.json
