#!/bin/bash
# Sweep the rows-per-group split of the forward / backward kernels on the small-Q workloads
# (PAVENET_MSDA_FWD_SPLIT / PAVENET_MSDA_BWD_SPLIT; 0 = the library's heuristic).
OUT=gpurun_out/${1:-sweep}
mkdir -p $OUT
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  for s in 0 2 4 8 16 32; do
    PAVENET_MSDA_FWD_SPLIT=$s PAVENET_MSDA_BWD_SPLIT=$s timeout 200 python bench.py --steps 200 --warmup 20 \
      --workload $wl --no-cpu-baseline --no-e2e 2>>$OUT/err.log > $OUT/${wl}_s$s.json
    python - <<PY
import json
try:
    d = json.load(open('$OUT/${wl}_s$s.json')); k = d['kernel_ms']
    print('%-14s split %2d  fwd %.4f  zero %.4f  bwd %.4f ms' % ('$wl', $s, k['fwd'], k['grad_value_zero_fill'], k['bwd']))
except Exception as e:
    print('$wl', $s, 'ERR', e)
PY
  done
done | tee $OUT/summary.txt
