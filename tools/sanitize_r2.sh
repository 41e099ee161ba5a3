#!/bin/bash
# Round 2: compute-sanitizer over the kernel-family tests (every kernel family and variant of the
# sampling op, small shapes): memcheck, then racecheck (shared-memory hazards: record boards, the
# staged windows of the tile kernel).  Full-size cases are not run under the sanitizer.
OUT=gpurun_out/${1:-sanitize_r2}
mkdir -p $OUT
SEL="tests/test_gpu_kernel_families.py"
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 300 \
      python -m pytest $SEL -m gpu -q -x -p no:cacheprovider > $OUT/$tool.log 2>&1
  echo "exit $?" >> $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " $OUT/$tool.log | tail -5
done
