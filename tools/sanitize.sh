#!/bin/bash
# compute-sanitizer passes over the kernels (memcheck, racecheck, initcheck on a small set of tests).
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
SEL='rows_kernels_match_oracle or fused_function_matches or host_buffer_entry_points or golden'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --target-processes all \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $OUT/$tool.log 2>&1
  echo "exit $?" | tee -a $OUT/$tool.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $OUT/$tool.log | tail -4
done
