#!/bin/bash
# Memory-safety evidence for the product library.
# 1. compute-sanitizer (memcheck / racecheck / synccheck / initcheck) on the small all-kernels workload
#    (tests/tools/sanitize_workload.py) -- where the pool allows it; its refusal is logged otherwise.
# 2. SWB_GUARD=1: every device arena between two 4 KiB canary zones and without over-allocation; the workload and the GPU
#    test suite check the zones after every stage / test (swb_debug_guard_check).
# Logs -> gpurun_out/sanitizer_*.log; the summaries are copied to profiles/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check full"
  [ $tool = racecheck ] && extra="--racecheck-report all"
  scale=1; [ $tool = racecheck ] && scale=${RACE_SCALE:-0.25}
  ( time SWB_SANITIZE_SCALE=$scale timeout ${SAN_TIMEOUT:-600} $CS --tool $tool $extra --error-exitcode 9 --print-limit 30 \
      python tests/tools/sanitize_workload.py ) > gpurun_out/sanitizer_${tool}_workload.log 2>&1
  echo "$tool workload: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done in|LEAK SUMMARY|closed" gpurun_out/sanitizer_${tool}_workload.log | tail -3
done
( time SWB_GUARD=1 python tests/tools/sanitize_workload.py ) > gpurun_out/guard_workload.log 2>&1
echo "guard workload: rc=$?"; tail -4 gpurun_out/guard_workload.log
( time SWB_GUARD=1 python -m pytest tests -q -m gpu -x ) > gpurun_out/guard_pytest_gpu.log 2>&1
echo "guard suite: rc=$?"; tail -4 gpurun_out/guard_pytest_gpu.log
