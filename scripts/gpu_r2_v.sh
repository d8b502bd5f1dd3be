#!/bin/bash
# round-2 call "v": final single-GPU state -- whole GPU suite, World bench (parity + breakdown), Twitter-US bench
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/v_smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/v_pytest_gpu.log 2>&1 ) 2> gpurun_out/v_pytest_gpu.time
echo "pytest gpu rc=$?"; tail -6 gpurun_out/v_pytest_gpu.log; tail -3 gpurun_out/v_pytest_gpu.time
timeout 900 python bench.py --breakdown > gpurun_out/v_bench_world.json 2> gpurun_out/v_bench_world.log; echo "bench world rc=$?"
grep -A16 "op breakdown" gpurun_out/v_bench_world.log | cut -c1-110
timeout 600 python bench.py --workload twitter-us > gpurun_out/v_bench_us.json 2> gpurun_out/v_bench_us.log; echo "bench us rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/v_bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    p = d.get("parity", {})
    print("%-28s value %8.3f e2e %8.3f frac %.4f loss %r driver %s" % (f.split("/")[-1], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["loss"], d.get("engine", d["config"]).get("epoch_driver", "")[:30]))
    print("    parity max %.3f (%s) over-noise %.3f over: %s" % (p.get("max_scaled_err", -1), p.get("worst_check"), p.get("max_scaled_err_over_reference_noise", -1), p.get("checks_over_tolerance")))
    print("    noise", p.get("reference_f32_noise"))
PY
