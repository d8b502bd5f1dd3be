#!/bin/bash
# round 2, call P (1 GPU): stacked X^T blocks + gate epilogue class: tests, X^T.dZ1 in isolation, epoch breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_layers.py tests/test_gpu_mlpconv.py tests/test_fit_golden.py tests/test_layers_golden.py -m gpu -x -q --timeout 600 > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/p_pytest.log
: > gpurun_out/p_xt_sweep.jsonl
for cfg in "GCG_X=1" "GCG_XT_STACKED=0" "GCG_XT_STACKED_KERNEL=gather" "GCG_XT_BLOCK_MB=64" "GCG_XT_HEAVY_FACTOR=8" "GCG_XT_HEAVY_FACTOR=4 GCG_XT_BLOCK_MB=64"; do
  env $cfg timeout 600 python scripts/xt_sweep.py >> gpurun_out/p_xt_sweep.jsonl 2>> gpurun_out/p_xt_sweep.err
  echo "xt [$cfg] rc=$?"; tail -1 gpurun_out/p_xt_sweep.jsonl | cut -c1-220
done
timeout 1200 python bench.py --workload twitter-world --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/p_bench_world.json 2> gpurun_out/p_bench_world.log
echo "bench rc=$?"; grep -A20 "op breakdown" gpurun_out/p_bench_world.log | cut -c1-120; grep "parity" gpurun_out/p_bench_world.log | awk '{ for(i=1;i<=NF;i++) if ($i=="scaled") v=$(i+1); print v, $0 }' | sort -n -r | head -3 | cut -c1-170; tail -1 gpurun_out/p_bench_world.log | cut -c1-200
