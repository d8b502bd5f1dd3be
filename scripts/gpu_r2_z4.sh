#!/bin/bash
# round-2 call "z4": cursor + next-span prefetch during the drain; same-box baseline (one CTA per 16 spans, span 384)
mkdir -p gpurun_out
GCG_STREAM_PERSISTENT=2 GCG_STREAM_SPAN=64 timeout 300 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_layers.py -m gpu -x -q --timeout 200 > gpurun_out/z4_pytest.log 2>&1; echo "pytest (cursor, span 64) rc=$?"; tail -2 gpurun_out/z4_pytest.log
timeout 300 python -m pytest tests/test_gpu_spmm.py -m gpu -x -q --timeout 200 > gpurun_out/z4_pytest_default.log 2>&1; echo "pytest (default) rc=$?"; tail -1 gpurun_out/z4_pytest_default.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-cpu-baseline --breakdown > gpurun_out/z4_$tag.json 2> gpurun_out/z4_$tag.log; echo "$tag rc=$?"; }
run base384 GCG_STREAM_PERSISTENT=0
run cursor128 GCG_STREAM_PERSISTENT=2 GCG_STREAM_SPAN=128
run cursor64 GCG_STREAM_PERSISTENT=2 GCG_STREAM_SPAN=64
run cursor32 GCG_STREAM_PERSISTENT=2 GCG_STREAM_SPAN=32
python - <<'PY'
import json
for tag in ("base384", "cursor128", "cursor64", "cursor32"):
    try:
        d = json.loads(open("gpurun_out/z4_%s.json" % tag).read().strip().splitlines()[-1])
        ops = {o["op"][:44]: round(o["ms"], 2) for o in d["breakdown"]["ops"] if o["op"].startswith("spmm")}
        print("%-10s epoch %.2f  A_hat.H %.3f ms (frac %.4f)  X.W1 %.2f  X^T.dZ1 %.2f  loss %r" % (tag, d["value"], d["roofline"]["ms"], d["roofline"]["frac"],
              d["roofline"]["other_sparse_products"][0]["ms"], d["roofline"]["other_sparse_products"][1]["ms"], d["loss"]))
        print("     ", ops)
    except Exception as e:
        print(tag, "unreadable", e)
PY
