#!/bin/bash
# A/B alternative builds of the library (PAVENET_MSDA_LIB) on the op workloads.
OUT=gpurun_out/${1:-ab}; shift
mkdir -p $OUT
for lib in "$@"; do
  for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1 encoder_cfg2; do
    PAVENET_MSDA_LIB=$lib timeout 200 python bench.py --steps 200 --warmup 20 --workload $wl \
        --no-cpu-baseline --no-e2e 2>>$OUT/err.log > $OUT/tmp.json
    python - <<PY
import json
try:
    d = json.load(open('$OUT/tmp.json')); k = d['kernel_ms']
    print('%-28s %-14s fwd %.4f  zero %.4f  bwd %.4f ms' % ('$lib'.split('/')[-1], '$wl', k['fwd'], k['grad_value_zero_fill'], k['bwd']))
except Exception as e:
    print('$lib', '$wl', 'ERR', e)
PY
  done
done | tee $OUT/summary.txt
