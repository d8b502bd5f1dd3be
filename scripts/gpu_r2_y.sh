#!/bin/bash
# round-2 call "y" (4 GPUs): the World bench at N=4 with the final code (feature slices + peer-memory transposes)
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/y_bench_g4.json 2> gpurun_out/y_bench_g4.log; echo "bench g4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/y_bench_g4.json").read().strip().splitlines()[-1])
p = d.get("parity", {})
print("g4 value %.3f e2e %.3f loss %r %s" % (d["value"], d["e2e"]["value"], d["loss"], d.get("engine")))
print("   parity max %.3f (%s) over: %s" % (p.get("max_scaled_err", -1), p.get("worst_check"), p.get("checks_over_tolerance")))
PY
tail -5 gpurun_out/y_bench_g4.log | cut -c1-200
