#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mlpconv.py tests/test_gpu_layers.py -q -m gpu --timeout 600 2>&1 | tail -5
for wl in "$@"; do
timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --breakdown > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.log
echo "bench $wl exit $?"; grep -A24 "op breakdown" gpurun_out/bench_$wl.log; grep "epoch" gpurun_out/bench_$wl.log
done
