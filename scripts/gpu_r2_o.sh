#!/bin/bash
# round 2, call O (8 GPUs): the epoch at 8 ranks with phase timings and the sampled-row parity block
mkdir -p gpurun_out
GCG_DIST_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 \
    bench.py --gpus 8 --steps 5 --warmup 3 --breakdown --parity-rows 256 > gpurun_out/o_bench_g8.json 2> gpurun_out/o_bench_g8.log
echo "bench g8 rc=$?"; grep -E "x[0-9]+ +[0-9.]+ ms$|epoch .* ms \(min" gpurun_out/o_bench_g8.log | head -10; grep -A16 "op breakdown" gpurun_out/o_bench_g8.log | cut -c1-110
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/o_bench_g8.json").read())
    p = d.get("parity") or {}
    print("g8", d["value"], "e2e", d["e2e"]["value"], "parity", p.get("max_scaled_err"), p.get("worst_check"), p.get("max_err_over_ref_max"), "loss", d["loss"], "parity s", p.get("seconds"))
except Exception as e:
    print("no json", e)
PY
grep -iE "error|Traceback" gpurun_out/o_bench_g8.log | head -5
