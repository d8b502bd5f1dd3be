#!/bin/bash
# round-2 call "r": whole GPU test suite, default bench (with parity + breakdown), bench with the opt-in
# single-pass TF32 rule for long weight-gradient contractions
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r_pytest_gpu.log 2>&1 ) 2> gpurun_out/r_pytest_gpu.time
echo "pytest gpu rc=$?"; tail -3 gpurun_out/r_pytest_gpu.log; tail -3 gpurun_out/r_pytest_gpu.time
timeout 900 python bench.py --breakdown > gpurun_out/r_bench_default.json 2> gpurun_out/r_bench_default.log
echo "bench default rc=$?"; grep -A14 "op breakdown" gpurun_out/r_bench_default.log | cut -c1-110
grep "parity L3\|parity L2 H.W\|parity L2 dW\|parity L3 dW" gpurun_out/r_bench_default.log | cut -c1-220
GCG_GEMM_LONGK_TF32=65536 timeout 900 python bench.py --breakdown > gpurun_out/r_bench_longk_tf32.json 2> gpurun_out/r_bench_longk_tf32.log
echo "bench longk rc=$?"; grep -A14 "op breakdown" gpurun_out/r_bench_longk_tf32.log | cut -c1-110
grep "parity L2 dW\|parity L3 dW\|parity L2 dWg" gpurun_out/r_bench_longk_tf32.log | cut -c1-220
python - <<'PY'
import json
for f in ("r_bench_default", "r_bench_longk_tf32"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        p = d.get("parity", {})
        print(f, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "frac", round(d["roofline"]["frac"], 4),
              "parity", round(p.get("max_scaled_err", -1), 3), p.get("worst_check"), "over noise", round(p.get("max_scaled_err_over_reference_noise", -1), 3))
        print("   noise", p.get("reference_f32_noise"))
    except Exception as e:
        print(f, "unreadable", e)
PY
