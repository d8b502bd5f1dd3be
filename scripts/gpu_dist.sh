#!/bin/bash
N=${1:-2}; WL=${2:-twitter-us}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist.py -q -m gpu -x --timeout 800 > gpurun_out/pytest_dist.log 2>&1
echo "pytest dist exit $?"; tail -15 gpurun_out/pytest_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL > gpurun_out/bench_${WL}_g$N.json 2> gpurun_out/bench_${WL}_g$N.log
echo "bench exit $?"; grep -v "^\[rank [1-9]" gpurun_out/bench_${WL}_g$N.log | tail -8; cat gpurun_out/bench_${WL}_g$N.json
