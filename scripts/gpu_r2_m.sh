#!/bin/bash
# round 2, call M (1 GPU): compute-sanitizer on the small configurations, then BASELINE config 5 sweep (part 1)
mkdir -p gpurun_out
bash scripts/gpu_r2_sanitize.sh
for n in 1000000 2000000; do
  timeout 1200 python scripts/spmm_sweep.py --n $n --deg 5 10 20 50 --F 64 128 256 300 600 --graph chunglu community-sorted --panel -9 0 >> gpurun_out/m_config5_sweep.jsonl 2>> gpurun_out/m_config5_sweep.err
  echo "sweep n=$n rc=$?"
done
wc -l gpurun_out/m_config5_sweep.jsonl
