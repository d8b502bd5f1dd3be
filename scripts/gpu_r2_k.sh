#!/bin/bash
# round 2, call K (8 GPUs): the epoch at 8 and 4 ranks with phase timings, then the P2P probe
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/k_topo.txt 2>&1
run() { # n, name, extra args
  n=$1; name=$2; shift 2
  GCG_DIST_PROFILE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $n --steps 8 --warmup 3 --breakdown "$@" > gpurun_out/k_bench_$name.json 2> gpurun_out/k_bench_$name.log
  echo "bench $name rc=$?"; grep -E "x[0-9]+ +[0-9.]+ ms$|epoch .* ms \(min" gpurun_out/k_bench_$name.log | head -10; grep -A12 "op breakdown" gpurun_out/k_bench_$name.log | cut -c1-110
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/k_bench_$name.json").read())
    print("$name", d["value"], "e2e", d["e2e"]["value"], "parity", (d.get("parity") or {}).get("max_scaled_err"), "loss", d["loss"])
except Exception as e:
    print("no json", e)
PY
}
run 8 g8 --parity-rows 512
run 4 g4 --no-parity
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 scripts/p2p_probe.py > gpurun_out/k_p2p_probe_8gpu.json 2> gpurun_out/k_p2p_probe_8gpu.err
echo "probe rc=$?"; tail -5 gpurun_out/k_p2p_probe_8gpu.json | cut -c1-300
