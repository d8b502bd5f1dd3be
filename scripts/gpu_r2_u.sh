#!/bin/bash
# round-2 call "u": determinism of the World epoch on one box, put_rows vs index_copy_, GEOTEXT configs 1/2, Twitter-US record
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/u_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/u_smoke.log
B="--no-parity --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/u_world_a.json 2> gpurun_out/u_world_a.log; echo "world a rc=$?"
timeout 600 python bench.py $B > gpurun_out/u_world_b.json 2> gpurun_out/u_world_b.log; echo "world b rc=$?"
timeout 600 python scripts/exp_putrows_torch.py $B > gpurun_out/u_world_torchput.json 2> gpurun_out/u_world_torchput.log; echo "world torch put rc=$?"
for v in graph nograph native_nograph; do
  case $v in
    graph) E=""; F="";;
    nograph) E=""; F="--no-graph";;
    native_nograph) E="GCG_NATIVE_EPOCH=1"; F="--no-graph";;
  esac
  env $E timeout 300 python bench.py --workload geotext --steps 50 --warmup 5 $F > gpurun_out/u_geotext_$v.json 2> gpurun_out/u_geotext_$v.log; echo "geotext $v rc=$?"
done
timeout 300 python bench.py --impl reference --workload geotext --steps 5 --warmup 1 > gpurun_out/u_geotext_reference_3layer.json 2> gpurun_out/u_geotext_reference_3layer.log; echo "geotext ref3 rc=$?"
timeout 300 python bench.py --impl reference --workload geotext --layers 2 --highway 0 --steps 5 --warmup 1 > gpurun_out/u_geotext_reference_2layer.json 2> gpurun_out/u_geotext_reference_2layer.log; echo "geotext ref2 rc=$?"
timeout 300 python bench.py --workload geotext --layers 2 --highway 0 --steps 50 --warmup 5 > gpurun_out/u_geotext_2layer.json 2> gpurun_out/u_geotext_2layer.log; echo "geotext 2layer rc=$?"
timeout 600 python bench.py --workload twitter-us > gpurun_out/u_us.json 2> gpurun_out/u_us.log; echo "us rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/u_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-48s value %10.3f e2e %10.3f loss %r acc %r %s" % (f.split("/")[-1], d["value"], d.get("e2e", {}).get("value", -1), d.get("loss"), d.get("acc"), d.get("engine", d["config"]).get("epoch_driver", "")[:40]))
        if "parity" in d:
            print("      parity", d["parity"]["max_scaled_err"], d["parity"]["worst_check"], d["parity"].get("max_scaled_err_over_reference_noise"))
    except Exception as e:
        print(f, "unreadable", e)
PY
