#!/bin/bash
# Experiment: encoder queries handed to the op in patch order instead of raster order.
OUT=gpurun_out/${1:-order}
mkdir -p $OUT
for o in raster patch8x4 patch4x8 patch16x2 patch8x8 patch16x8 patch16x16 patch32x32; do
  PAVENET_BENCH_QUERY_ORDER=$o timeout 200 python bench.py --steps 100 --warmup 10 \
      --workload encoder_cfg2 --no-cpu-baseline --no-e2e 2>>$OUT/err.log > $OUT/$o.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$o.json')); k = d['kernel_ms']
    print('%-12s fwd %.4f  bwd %.4f ms' % ('$o', k['fwd'], k['bwd']))
except Exception as e:
    print('$o', 'ERR', e)
PY
done | tee $OUT/summary.txt
