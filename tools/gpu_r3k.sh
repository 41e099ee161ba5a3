#!/bin/bash
# fused-prologue backward (rows family): 2 resident blocks per SM at 128 registers vs 3 at 85 with spills
OUT=gpurun_out/r3k
mkdir -p $OUT
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --fused"
for lib in libpavenet_msda.so libpavenet_msda_fb3.so libpavenet_msda.so libpavenet_msda_fb3.so; do
  PAVENET_MSDA_LIB=$PWD/pavenet_b200/lib/$lib timeout 300 $B > $OUT/tmp.json 2>>$OUT/err.log
  python - <<PY
import json
d = json.load(open('$OUT/tmp.json')); k = d['kernel_ms']
print('%-28s fused encoder_cfg2: fwd %.4f zero %.4f bwd %.4f step %.4f' % ('$lib', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step']))
PY
done
tail -2 $OUT/err.log
