#!/bin/bash
# round 2, call I (1 GPU): dense head / sparse tail split of X -- tests, World + US epochs with breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_mlpconv.py tests/test_gpu_parity_full_size.py -m gpu -x -q --timeout 600 > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/i_pytest.log
for wl in twitter-world twitter-us; do
timeout 1200 python bench.py --workload $wl --steps 5 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/i_bench_$wl.json 2> gpurun_out/i_bench_$wl.log
echo "bench $wl rc=$?"; grep -A14 "op breakdown" gpurun_out/i_bench_$wl.log | cut -c1-120; grep "parity" gpurun_out/i_bench_$wl.log | awk '{ for(i=1;i<=NF;i++) if ($i=="scaled") v=$(i+1); print v, $0 }' | sort -n -r | head -4 | cut -c1-170; tail -1 gpurun_out/i_bench_$wl.log | cut -c1-200
done
