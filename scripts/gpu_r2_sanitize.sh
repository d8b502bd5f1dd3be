#!/bin/bash
# compute-sanitizer on the small configurations: memcheck for every hand-written kernel family the tests reach,
# racecheck for the kernels that synchronise through shared memory / mbarriers (bulk-copy SpMM, tcgen05 GEMM).
mkdir -p gpurun_out
CS="compute-sanitizer --error-exitcode 9 --print-limit 20"
run() { # tag, tool, pytest args...
  tag=$1; tool=$2; shift 2
  timeout 1500 $CS --tool $tool python -m pytest "$@" -x -q -p no:cacheprovider > gpurun_out/san_${tool}_$tag.log 2>&1
  echo "$tool $tag rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/san_${tool}_$tag.log | tail -3
}
run spmm_stream memcheck tests/test_gpu_spmm.py -k "streaming or bit_exact_vs_scipy or hub_rows"
run gemm memcheck tests/test_gpu_gemm.py
run layers memcheck tests/test_gpu_layers.py tests/test_gpu_elementwise.py
run model memcheck tests/test_gpu_mlpconv.py -k "one_training_step or cuda_graph or tensor_core"
run spmm_bulk racecheck tests/test_gpu_spmm.py -k "bit_exact_vs_scipy and 600"
run gemm racecheck tests/test_gpu_gemm.py -k "tensor_core or presplit"
