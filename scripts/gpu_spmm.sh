#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/spmm_sweep.py --n 450000 --deg 20 --F 600 256 64 --graph community --reorder community --panel 0 --tune 0,0 2,0 2,2 4,0 2,4 0,0 2,0 > gpurun_out/sweep_ab.jsonl 2> gpurun_out/sweep_ab.log
echo "sweep exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/sweep_ab.jsonl'):
    d=json.loads(l)
    print("%-9s ro=%-9s F=%-4d panel=%-3d tune=%-4s %.3f ms (min %.3f)  alg %.0f GB/s (%.3f)  gather %.0f GB/s"%(d['graph'],d['reorder'],d['F'],d['panel'],d['tune'],d['ms'],d['ms_min'],d['alg_GBps'],d['frac_of_measured_peak'],d['gather_GBps']))
PY
tail -3 gpurun_out/sweep_ab.log
