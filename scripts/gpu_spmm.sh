#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmm.py -q -m gpu -x --timeout 300 2>&1 | tail -6
timeout 600 python scripts/spmm_sweep.py --n 450000 --deg 20 --F 64 256 600 1024 --graph chunglu community --reorder none community --panel 0 -1 > gpurun_out/sweep_tma.jsonl 2> gpurun_out/sweep_tma.log
echo "sweep exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/sweep_tma.jsonl'):
    d=json.loads(l); print("%-9s ro=%-9s F=%-4d panel=%-3d  %.3f ms  alg %.0f GB/s (%.3f)  gather %.0f GB/s"%(d['graph'],d['reorder'],d['F'],d['panel'],d['ms'],d['alg_GBps'],d['frac_of_measured_peak'],d['gather_GBps']))
PY
tail -3 gpurun_out/sweep_tma.log
