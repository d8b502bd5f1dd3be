#!/bin/bash
# round-2 call "w" (2 GPUs): NCCL world-2 parity cases (torch.distributed and native C-ABI collectives), World bench at N=2 both ways
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q --timeout 800 > gpurun_out/w_pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -15 gpurun_out/w_pytest_dist.log | cut -c1-200
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/w_bench_g2_$tag.json 2> gpurun_out/w_bench_g2_$tag.log; echo "bench g2 $tag rc=$?"; }
run torch GCG_DIST_COMM=torch
run native GCG_DIST_COMM=native
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/w_bench_g2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        p = d.get("parity", {})
        print("%-28s value %8.3f e2e %8.3f loss %r %s | %s" % (f.split("/")[-1], d["value"], d["e2e"]["value"], d["loss"], d.get("engine", d["config"]).get("parallelism"), d.get("engine", d["config"]).get("collectives")))
        print("    parity max %.3f (%s) over-noise %.3f" % (p.get("max_scaled_err", -1), p.get("worst_check"), p.get("max_scaled_err_over_reference_noise", -1)))
    except Exception as e:
        print(f, "unreadable", e)
PY
