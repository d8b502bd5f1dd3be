#!/bin/bash
# round 2, call G (1 GPU): whole GPU test suite; GEMM chain length A/B (rates + World parity)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/g_pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -6 gpurun_out/g_pytest_all.log
for ck in 4 2; do
  GCG_GEMM_CHUNK_KB=$ck timeout 600 python scripts/tc_check.py > gpurun_out/g_tc_check_ck$ck.log 2>&1; echo "tc_check ck=$ck rc=$?"; grep "bench tf32x3" gpurun_out/g_tc_check_ck$ck.log
  GCG_GEMM_CHUNK_KB=$ck timeout 900 python bench.py --workload twitter-world --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_world_ck$ck.json 2> gpurun_out/g_bench_world_ck$ck.log
  echo "bench ck=$ck rc=$?"; grep "parity" gpurun_out/g_bench_world_ck$ck.log | awk '{ for(i=1;i<=NF;i++) if ($i=="scaled") v=$(i+1); print v, $0 }' | sort -n -r | head -3 | cut -c1-170; tail -1 gpurun_out/g_bench_world_ck$ck.log | cut -c1-200
done
