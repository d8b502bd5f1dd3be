#!/bin/bash
N=${1:-2}; WL=${2:-twitter-us}
mkdir -p gpurun_out
if [ "$3" != "notest" ]; then timeout 600 python -m pytest tests/test_dist.py -q -m gpu -x --timeout 500 > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist exit $?"; tail -25 gpurun_out/pytest_dist.log | grep -E "dist case|passed|failed|Error|error" | head; fi
for mode in "--partition feature" "--partition feature --no-peer-memory" "--partition row"; do
tag=$(echo $mode | tr -d ' -'); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL $mode > gpurun_out/bench_${WL}_g${N}_$tag.json 2> gpurun_out/bench_${WL}_g${N}_$tag.log
echo "bench $mode exit $?"; grep -E "Error|error|Traceback" gpurun_out/bench_${WL}_g${N}_$tag.log | head -3; python -c "
import json; d=json.load(open('gpurun_out/bench_${WL}_g${N}_$tag.json')); print(d['value'], d['config']['parallelism'], d['roofline']['ms'])"
done
