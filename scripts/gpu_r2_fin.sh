#!/bin/bash
# round-2 final call: whole GPU suite + the World bench (parity + breakdown) with the final defaults
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/fin_pytest_gpu.log 2>&1 ) 2> gpurun_out/fin_pytest_gpu.time
echo "pytest gpu rc=$?"; tail -4 gpurun_out/fin_pytest_gpu.log; tail -3 gpurun_out/fin_pytest_gpu.time
timeout 600 python bench.py --breakdown > gpurun_out/fin_bench_world.json 2> gpurun_out/fin_bench_world.log; echo "bench world rc=$?"
grep -A14 "op breakdown" gpurun_out/fin_bench_world.log | cut -c1-110
python - <<'PY'
import json
d = json.loads(open("gpurun_out/fin_bench_world.json").read().strip().splitlines()[-1])
p = d.get("parity", {})
print("value %.3f e2e %.3f A_hat.H %.3f ms frac %.4f loss %r engine %s" % (d["value"], d["e2e"]["value"], d["roofline"]["ms"], d["roofline"]["frac"], d["loss"], d.get("engine")))
print("parity max %.3f (%s) over: %s" % (p.get("max_scaled_err", -1), p.get("worst_check"), p.get("checks_over_tolerance")))
print("noise", p.get("reference_f32_noise"))
PY
