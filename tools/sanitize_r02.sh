#!/bin/bash
# compute-sanitizer passes over the kernels added in round 2 (memcheck + racecheck), logs under gpurun_out/
mkdir -p gpurun_out
SEL='rows or linear_tc or lookup or corr_pyramid or gru_pass or stem or split_residual'
for tool in memcheck racecheck; do
  [ "$tool" = racecheck ] && SEL="rows or linear_tc or stem or split_residual"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_kernels.py tests/test_gpu_decoder.py -x -q -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_$tool.log | tail -4
done
