#!/bin/bash
# compute-sanitizer over the kernels added late in round 1: TMA-store GEMM epilogue (128- and 256-row CTAs),
# LayerNorm, fused FFN, bf16-value backward with 8 lanes per row.
OUT=gpurun_out/${1:-san2}
mkdir -p $OUT
SEL='linear256_matches_fp64 or linear256_masks or bf16_value_storage or layer_norm or fused_ffn_matches'
run() {  # name tool [env...]
  name=$1; tool=$2; shift 2
  echo "== $name ($tool)"
  sel="$SEL"
  # racecheck slows kernels down by orders of magnitude: leave the 66 669-row cases to memcheck
  [ $tool = racecheck ] && sel="($SEL) and not 66669"
  env "$@" timeout 300 compute-sanitizer --tool $tool --error-exitcode 99 --target-processes all \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel" > $OUT/$name.log 2>&1
  echo "exit $?" | tee -a $OUT/$name.log
  grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" $OUT/$name.log | tail -3
}
run memcheck memcheck A=1
run memcheck_bm256 memcheck PAVENET_MSDA_LINEAR_BM=256
run racecheck racecheck A=1
