#!/bin/bash
# Round-end style validation on one B200: tests, smoke, the three bench workloads, reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/BENCH_world.json 2> gpurun_out/BENCH_world.log; echo "bench world exit $?"; grep "epoch" gpurun_out/BENCH_world.log
timeout 600 python bench.py --workload twitter-us --breakdown > gpurun_out/BENCH_us.json 2> gpurun_out/BENCH_us.log; echo "bench us exit $?"; grep "epoch" gpurun_out/BENCH_us.log
timeout 600 python bench.py --workload geotext --breakdown > gpurun_out/BENCH_geotext.json 2> gpurun_out/BENCH_geotext.log; echo "bench geotext exit $?"; grep "epoch" gpurun_out/BENCH_geotext.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/BENCH_ref_world.json 2> gpurun_out/BENCH_ref_world.log; echo "ref exit $?"
python - <<'PY'
import json
for n in ("world","us","geotext","ref_world"):
    try:
        d=json.load(open("gpurun_out/BENCH_%s.json"%n))
        print(n, "value %.2f ms"%d["value"], "e2e", d.get("e2e",{}).get("value"), "roof", {k:d["roofline"][k] for k in ("achieved","frac","ms")} if "roofline" in d else None, "cpu", d.get("cpu_baseline",{}).get("value"), "launches", d.get("launches_per_epoch"))
    except Exception as e: print(n, "ERR", e)
PY
