#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_world.json 2> gpurun_out/bench_world.log
echo "bench world exit $?"; tail -4 gpurun_out/bench_world.log; cat gpurun_out/bench_world.json
timeout 900 python scripts/spmm_sweep.py --n 450000 --deg 20 --F 64 600 --graph chunglu community --reorder none community degree --panel 0 16 32 > gpurun_out/sweep_us.jsonl 2> gpurun_out/sweep_us.log
echo "sweep exit $?"; cat gpurun_out/sweep_us.jsonl; tail -3 gpurun_out/sweep_us.log
