#!/bin/bash
# round 2, call H (2 GPUs): NCCL world-2 parity of every multi-GPU scheme, then the feature-sliced epoch
# with / without the fused second transpose and the flag barrier (phase timings on rank 0)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/h_pytest_dist.log 2>&1
echo "pytest dist rc=$?"; tail -4 gpurun_out/h_pytest_dist.log
run() { # name, env...
  name=$1; shift
  env "$@" GCG_DIST_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 5 --warmup 3 --partition feature --no-parity > gpurun_out/h_bench_$name.json 2> gpurun_out/h_bench_$name.log
  echo "bench $name rc=$?"; grep -E "x[0-9]+ +[0-9.]+ ms$|epoch .* ms \(min" gpurun_out/h_bench_$name.log | head -12
}
run fused_flag GCG_X=1
run unfused_flag GCG_DIST_FUSED_PUSH=0
run unfused_nccl GCG_DIST_FUSED_PUSH=0 GCG_DIST_NCCL_BARRIER=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
      bench.py --gpus 2 --steps 5 --warmup 3 --parity-rows 512 > gpurun_out/h_bench_auto.json 2> gpurun_out/h_bench_auto.log
echo "bench auto rc=$?"; tail -3 gpurun_out/h_bench_auto.log | cut -c1-300; cut -c1-600 gpurun_out/h_bench_auto.json
