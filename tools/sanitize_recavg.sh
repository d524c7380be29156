#!/bin/bash
# compute-sanitizer over the one-launch RecAvg backward and the tensor-core path helpers (small cases: the tools slow kernels 10-100x)
mkdir -p gpurun_out
K='recavg_bwd_fused_equals and (5-6-7-64 or 3-1-1-8 or 2-5-24-776 or 9-30-24-256)'
for tool in memcheck racecheck; do
  timeout 170 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -8
done
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_full.py -q -x -k "tensor_core_path and (5-7-64 or 70-40)" > gpurun_out/r2_sanitizer_memcheck_tc.log 2>&1
echo "== memcheck tc rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_memcheck_tc.log | head -4
