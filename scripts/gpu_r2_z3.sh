#!/bin/bash
# round-2 call "z3": persistent grid with spans handed out by an atomic cursor x span length
mkdir -p gpurun_out
GCG_STREAM_PERSISTENT=2 GCG_STREAM_SPAN=64 timeout 300 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_layers.py -m gpu -x -q --timeout 200 > gpurun_out/z3_pytest_spmm_cursor.log 2>&1; echo "pytest spmm (cursor, span 64) rc=$?"; tail -2 gpurun_out/z3_pytest_spmm_cursor.log
for sp in 384 128 64; do
  GCG_STREAM_PERSISTENT=2 GCG_STREAM_SPAN=$sp timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-cpu-baseline --breakdown > gpurun_out/z3_cursor_span_$sp.json 2> gpurun_out/z3_cursor_span_$sp.log
  echo "cursor span $sp rc=$?"
done
python - <<'PY'
import json
for sp in (384, 128, 64):
    try:
        d = json.loads(open("gpurun_out/z3_cursor_span_%d.json" % sp).read().strip().splitlines()[-1])
        ops = {o["op"][:44]: round(o["ms"], 2) for o in d["breakdown"]["ops"] if o["op"].startswith("spmm")}
        print("cursor span %3d: epoch %.2f  A_hat.H %.3f ms (frac %.4f)  X.W1 %.2f  X^T.dZ1 %.2f  loss %r" % (sp, d["value"], d["roofline"]["ms"], d["roofline"]["frac"],
              d["roofline"]["other_sparse_products"][0]["ms"], d["roofline"]["other_sparse_products"][1]["ms"], d["loss"]))
        print("     ", ops)
    except Exception as e:
        print(sp, "unreadable", e)
PY
