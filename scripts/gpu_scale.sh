#!/bin/bash
N=${1:-8}; WL=${2:-twitter-world}; shift; shift
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload $WL --breakdown "$@" > gpurun_out/bench_${WL}_g$N.json 2> gpurun_out/bench_${WL}_g$N.log
echo "bench exit $?"; grep -A10 "op breakdown" gpurun_out/bench_${WL}_g$N.log | head -12; grep -A8 "phases of the" gpurun_out/bench_${WL}_g$N.log; grep -E "Error|Traceback" gpurun_out/bench_${WL}_g$N.log | head -3
python -c "
import json; d=json.load(open('gpurun_out/bench_${WL}_g$N.json')); print('RESULT', d['n_gpus'], d['value'], d['config']['parallelism'], d['roofline']['ms'])"
