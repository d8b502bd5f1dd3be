#!/bin/bash
# round-2 call "s": whole GPU suite (no -x), then the World bench driven by the native gcg_epoch program
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/s_pytest_gpu.log 2>&1 ) 2> gpurun_out/s_pytest_gpu.time
echo "pytest gpu rc=$?"; tail -8 gpurun_out/s_pytest_gpu.log; tail -3 gpurun_out/s_pytest_gpu.time
GCG_NATIVE_EPOCH=1 timeout 900 python bench.py --no-parity > gpurun_out/s_bench_native.json 2> gpurun_out/s_bench_native.log
echo "bench native rc=$?"; tail -3 gpurun_out/s_bench_native.log | cut -c1-200
python - <<'PY'
import json
a = json.loads(open("gpurun_out/s_bench_native.json").read().strip().splitlines()[-1])
b = json.loads(open("profiles/r02_bench_default_world.json").read().strip().splitlines()[-1])
print("native: value %.2f e2e %.2f loss %r acc %r driver %s" % (a["value"], a["e2e"]["value"], a["loss"], a["acc"], a.get("engine", a["config"]).get("epoch_driver")))
print("python: value %.2f e2e %.2f loss %r acc %r" % (b["value"], b["e2e"]["value"], b["loss"], b["acc"]))
print("same loss/acc bits:", a["loss"] == b["loss"] and a["acc"] == b["acc"], "launches", a["launches_per_epoch"], b["launches_per_epoch"])
PY
