#!/bin/bash
# round 2, call J (1 GPU): X^T.dZ1 knobs in isolation; narrow-operand SpMM lane groups; epoch with selective chaining
mkdir -p gpurun_out
: > gpurun_out/j_xt_sweep.jsonl
for cfg in "GCG_X=1" "GCG_XT_BLOCK_KERNEL=stream" "GCG_XT_HEAVY_FACTOR=16" "GCG_XT_HEAVY_FACTOR=32" "GCG_XT_HEAVY_FACTOR=16 GCG_XT_BLOCK_KERNEL=stream" \
           "GCG_XT_BLOCK_MB=96 GCG_XT_HEAVY_FACTOR=16" "GCG_XT_BLOCK_MB=128 GCG_XT_HEAVY_FACTOR=32" "GCG_X_HEAD=512" "GCG_X_HEAD=128" "GCG_X_HEAD=0"; do
  env $cfg timeout 600 python scripts/xt_sweep.py >> gpurun_out/j_xt_sweep.jsonl 2>> gpurun_out/j_xt_sweep.err
  echo "xt [$cfg] rc=$?"; tail -1 gpurun_out/j_xt_sweep.jsonl | cut -c1-250
done
timeout 600 python scripts/spmm_narrow_sweep.py > gpurun_out/j_narrow.jsonl 2> gpurun_out/j_narrow.err; echo "narrow rc=$?"; cat gpurun_out/j_narrow.jsonl
timeout 1200 python bench.py --workload twitter-world --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/j_bench_world.json 2> gpurun_out/j_bench_world.log
echo "bench rc=$?"; grep -A22 "op breakdown" gpurun_out/j_bench_world.log | cut -c1-120; grep "parity" gpurun_out/j_bench_world.log | awk '{ for(i=1;i<=NF;i++) if ($i=="scaled") v=$(i+1); print v, $0 }' | sort -n -r | head -3 | cut -c1-170; tail -1 gpurun_out/j_bench_world.log | cut -c1-200
