#!/bin/bash
# kernel timings of the op workloads only (no tests): tools/quick_bench.sh TAG [bench args...]
TAG=${1:-qb}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for wl in encoder_cfg2 pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  timeout 300 python bench.py --steps 200 --warmup 20 --workload $wl --no-cpu-baseline --no-e2e "$@" 2>>$OUT/err.log > $OUT/$wl.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$wl.json')); k = d['kernel_ms']
    print('%-14s step %.4f  fwd %.4f  zero %.4f  bwd %.4f ms' % ('$wl', d['ms_per_step'], k['fwd'], k['grad_value_zero_fill'], k['bwd']))
except Exception as e:
    print('$wl', 'ERR', e)
PY
done | tee $OUT/summary.txt
