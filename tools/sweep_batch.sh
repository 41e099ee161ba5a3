#!/bin/bash
# BASELINE config 5: batch sweep of the encoder stress shape until time per query flattens.
OUT=gpurun_out/${1:-batch}
mkdir -p $OUT
for f in 1 2 4 8 16 32; do
  timeout 300 python bench.py --steps 50 --warmup 5 --workload stress_cfg5 --frames $f --sets 3 \
      --no-cpu-baseline --no-e2e 2>>$OUT/err.log > $OUT/frames_$f.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/frames_$f.json')); k = d['kernel_ms']
    print('frames %2d  %.4g queries/s  step %.4f ms  fwd %.4f bwd %.4f  ns/query %.2f  hbm frac (step) %.3f' % (
        $f, d['value'], d['ms_per_step'], k['fwd'], k['bwd'], 1e6 * d['ms_per_step'] / d['config']['queries_per_step'], d['roofline_step']['frac']))
except Exception as e:
    print('frames', $f, 'ERR', e)
PY
done | tee $OUT/summary.txt
