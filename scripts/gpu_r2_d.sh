#!/bin/bash
# round 2, call D: more-warps ring variants x near windows; other widths; then the epoch with the streaming kernel on
mkdir -p gpurun_out
timeout 900 python scripts/spmm_stream_sweep.py --workload twitter-world --F 600 --spans 256 384 --variants 1 2 3 4 8 9 --near 0 65536 262144 --rows-only > gpurun_out/d_sweep_F600.jsonl 2> gpurun_out/d_sweep_F600.err
echo "sweep600 rc=$?"
timeout 900 python scripts/spmm_stream_sweep.py --workload twitter-world --F 76 256 1024 --spans 384 --variants 1 2 3 5 --near 0 --rows-only > gpurun_out/d_sweep_other.jsonl 2> gpurun_out/d_sweep_other.err
echo "sweep other rc=$?"
timeout 1200 python bench.py --workload twitter-world --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/d_bench_world.json 2> gpurun_out/d_bench_world.log
echo "bench rc=$?"; grep -A40 "op breakdown" gpurun_out/d_bench_world.log | head -60; tail -5 gpurun_out/d_bench_world.log
