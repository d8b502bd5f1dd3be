#!/bin/bash
# round 2, call C: parity tests for the compact streaming kernels + sampled full-size parity, then the sweep
mkdir -p gpurun_out
free -g > gpurun_out/c_host.txt; nproc >> gpurun_out/c_host.txt
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_layers_golden.py -m gpu -x -q > gpurun_out/c_pytest_spmm.log 2>&1
echo "pytest spmm rc=$?"; tail -3 gpurun_out/c_pytest_spmm.log
timeout 1500 python -m pytest tests/test_gpu_parity_full_size.py -m gpu -q -s > gpurun_out/c_pytest_parity.log 2>&1
echo "pytest parity rc=$?"; grep -E "passed|failed|scaled +[0-9]+\.[0-9]+ " gpurun_out/c_pytest_parity.log | awk '{ if ($0 ~ /scaled/) { if ($5+0 > 0.5) print } else print }' | tail -60
timeout 1200 python scripts/spmm_stream_sweep.py --workload twitter-world --F 600 --spans 256 --variants 1 2 3 4 5 7 > gpurun_out/c_stream_sweep_world.jsonl 2> gpurun_out/c_stream_sweep_world.err
echo "sweep rc=$?"; tail -3 gpurun_out/c_stream_sweep_world.err
