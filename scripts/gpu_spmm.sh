#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/spmm_sweep.py --n 450000 --deg 20 --F 600 --graph chunglu community --reorder none community --panel 0 384 256 128 --tune 0,0 2,2 2,3 3,2 3,3 4,2 4,3 6,2 2,4 > gpurun_out/sweep_tune.jsonl 2> gpurun_out/sweep_tune.log
echo "sweep exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/sweep_tune.jsonl'):
    d=json.loads(l)
    if d['graph']=='community' and d['reorder']=='none': continue
    print("%-9s ro=%-9s F=%-4d panel=%-3d tune=%-4s %.3f ms  alg %.0f GB/s (%.3f)  gather %.0f GB/s"%(d['graph'],d['reorder'],d['F'],d['panel'],d['tune'],d['ms'],d['alg_GBps'],d['frac_of_measured_peak'],d['gather_GBps']))
PY
tail -3 gpurun_out/sweep_tune.log
