#!/bin/bash
# round-2 call "z": span length of the streaming SpMM (non-zeros per warp) -- the LRU model predicts less L2 thrash
# with shorter spans; World bench (short, with per-op breakdown) per setting
mkdir -p gpurun_out
for sp in 384 192 128 96 64 32; do
  GCG_STREAM_SPAN=$sp timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-cpu-baseline --breakdown > gpurun_out/z_span_$sp.json 2> gpurun_out/z_span_$sp.log
  echo "span $sp rc=$?"
done
python - <<'PY'
import json
for sp in (384, 192, 128, 96, 64, 32):
    try:
        d = json.loads(open("gpurun_out/z_span_%d.json" % sp).read().strip().splitlines()[-1])
        ops = {o["op"][:44]: round(o["ms"], 2) for o in d["breakdown"]["ops"] if o["op"].startswith("spmm")}
        print("span %3d: epoch %.2f  A_hat.H %.3f ms (frac %.4f)  X.W1 %.2f  X^T.dZ1 %.2f  loss %r" % (sp, d["value"], d["roofline"]["ms"], d["roofline"]["frac"],
              d["roofline"]["other_sparse_products"][0]["ms"], d["roofline"]["other_sparse_products"][1]["ms"], d["loss"]))
        print("     ", ops)
    except Exception as e:
        print(sp, "unreadable", e)
PY
