#!/bin/bash
# round 2, call L (2 GPUs): clustered GEMM tests + rates; NCCL world-2 parity incl. forced Twitter-scale paths;
# feature-sliced and row-partitioned epochs with the sampled-row parity check switched on
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q --timeout 300 > gpurun_out/l_pytest_gemm.log 2>&1
echo "pytest gemm rc=$?"; tail -4 gpurun_out/l_pytest_gemm.log
timeout 600 python scripts/tc_check.py > gpurun_out/l_tc_check.log 2>&1; echo "tc_check rc=$?"; grep "bench tf32" gpurun_out/l_tc_check.log
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/l_pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -4 gpurun_out/l_pytest_dist.log | cut -c1-400
for part in feature auto; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
      bench.py --gpus 2 --steps 5 --warmup 3 --partition $part --parity-rows 512 > gpurun_out/l_bench_$part.json 2> gpurun_out/l_bench_$part.log
  echo "bench $part rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/l_bench_$part.json").read())
    p = d.get("parity") or {}
    print("$part", d["value"], "e2e", d["e2e"]["value"], "parity", p.get("max_scaled_err"), p.get("worst_check"), "rel", p.get("max_err_over_ref_max"), p.get("worst_relative_check"), "loss", d["loss"])
except Exception as e:
    print("no json", e)
PY
done
