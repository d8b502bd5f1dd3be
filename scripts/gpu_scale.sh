#!/bin/bash
N=${1:-8}; WL=${2:-twitter-world}
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL --breakdown > gpurun_out/bench_${WL}_g$N.json 2> gpurun_out/bench_${WL}_g$N.log
echo "bench exit $?"; grep -A22 "op breakdown" gpurun_out/bench_${WL}_g$N.log | head -24; grep -E "rank 0\] epoch|Error" gpurun_out/bench_${WL}_g$N.log | tail -3
