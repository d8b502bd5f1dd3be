#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmm.py -q -m gpu -x -k "blocked or hub or accumulate" 2>&1 | tail -3
for mb in 64 96 128 192; do
GCG_XT_BLOCK_MB=$mb timeout 600 python bench.py --workload ${1:-twitter-world} --steps 3 --warmup 3 --no-cpu-baseline --breakdown 2>&1 >/dev/null | grep -E "spmm 500000x1400000|spmm 250000x450000|rank 0\] epoch|prepared" | sed "s/^/mb=$mb /"
done
