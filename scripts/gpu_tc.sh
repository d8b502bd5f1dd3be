#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1
echo "tc_check exit $?"; cat gpurun_out/tc_check.log | tail -80
