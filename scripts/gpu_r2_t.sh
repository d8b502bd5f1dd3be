#!/bin/bash
# round-2 call "t": whole GPU suite, default bench (parity + breakdown), the same bench driven by gcg_epoch_run
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/t_pytest_gpu.log 2>&1 ) 2> gpurun_out/t_pytest_gpu.time
echo "pytest gpu rc=$?"; tail -6 gpurun_out/t_pytest_gpu.log; tail -3 gpurun_out/t_pytest_gpu.time
timeout 900 python bench.py --breakdown > gpurun_out/t_bench_default.json 2> gpurun_out/t_bench_default.log
echo "bench default rc=$?"
GCG_NATIVE_EPOCH=1 timeout 900 python bench.py --no-parity > gpurun_out/t_bench_native.json 2> gpurun_out/t_bench_native.log
echo "bench native rc=$?"
python - <<'PY'
import json
a = json.loads(open("gpurun_out/t_bench_native.json").read().strip().splitlines()[-1])
b = json.loads(open("gpurun_out/t_bench_default.json").read().strip().splitlines()[-1])
print("native: value %.2f e2e %.2f loss %r acc %r driver %s" % (a["value"], a["e2e"]["value"], a["loss"], a["acc"], a.get("engine", a["config"]).get("epoch_driver")))
print("python: value %.2f e2e %.2f loss %r acc %r driver %s" % (b["value"], b["e2e"]["value"], b["loss"], b["acc"], b.get("engine", b["config"]).get("epoch_driver")))
print("same loss/acc bits:", a["loss"] == b["loss"] and a["acc"] == b["acc"], "launches", a["launches_per_epoch"], b["launches_per_epoch"])
p = b.get("parity", {})
print("parity", p.get("max_scaled_err"), p.get("worst_check"), p.get("checks_over_tolerance"), p.get("reference_f32_noise"))
PY
