#!/bin/bash
N=${1:-2}; WL=${2:-twitter-us}
mkdir -p gpurun_out
if [ "$3" != "notest" ]; then timeout 600 python -m pytest tests/test_dist.py -q -m gpu -x --timeout 500 > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist exit $?"; tail -3 gpurun_out/pytest_dist.log; fi
for part in feature row; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL --partition $part > gpurun_out/bench_${WL}_g${N}_$part.json 2> gpurun_out/bench_${WL}_g${N}_$part.log
echo "bench $part exit $?"; grep -E "rank 0\] epoch|Error" gpurun_out/bench_${WL}_g${N}_$part.log | tail -3; python -c "
import json; d=json.load(open('gpurun_out/bench_${WL}_g${N}_$part.json')); print(d['value'], d['config']['parallelism'], d['roofline']['ms'])"
done
