#!/bin/bash
OUT=gpurun_out/r3m
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernel_families.py tests/test_gpu_full_size.py tests/test_gpu_reference_kernel.py -m gpu -q --tb=short -x 2>&1 | tail -3
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
for wl in encoder_cfg2 encoder_cfg2 pose_cfg3 pose_cfg3_t3 petr_cfg1 stress_cfg5; do
  timeout 300 $B --workload $wl > $OUT/$wl.json 2>>$OUT/err.log
  python - <<PY
import json
d = json.load(open('$OUT/$wl.json')); k = d['kernel_ms']
print('%-16s fwd %.4f zero %.4f bwd %.4f step %.4f (eager %.4f)' % ('$wl', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['ms_per_step_eager']))
PY
done
tail -2 $OUT/err.log
