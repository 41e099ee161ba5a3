#!/bin/bash
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernel_families.py -m gpu -q --tb=short -x 2>&1 | tail -5 | tee $OUT/pytest.log
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-34s fwd %.4f  zero %.4f  bwd %.4f  step %.4f ms   frac step %.3f' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['roofline_step']['frac']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  run ${wl}_base $wl
  run ${wl}_phase $wl --option flat_order=1
  run ${wl}_phase_sep $wl --option flat_order=1 --fold-clear 0
done
tail -3 $OUT/err.log
