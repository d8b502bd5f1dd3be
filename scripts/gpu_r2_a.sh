#!/bin/bash
# round 2, call A: new streaming SpMM -- parity tests, then the variant sweep at Twitter-World shape
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_layers_golden.py -m gpu -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 1200 python scripts/spmm_stream_sweep.py --workload twitter-world --F 600 --spans 256 512 > gpurun_out/a_stream_sweep_world.jsonl 2> gpurun_out/a_stream_sweep_world.err
echo "sweep rc=$?"
tail -3 gpurun_out/a_stream_sweep_world.err
