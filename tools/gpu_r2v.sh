#!/bin/bash
# (batch, blocks/SM) sweep of the flat kernels again, now with the phase-aligned piece order
OUT=gpurun_out/r2v
mkdir -p $OUT
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-28s fwd %.4f zero %.4f bwd %.4f | step %.4f ms (eager %.4f) | frac step %.3f' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['ms_per_step_eager'], d['roofline_step']['frac']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
for wl in pose_cfg3 pose_cfg3_t3; do
  run ${wl}_base $wl
  for c in 1 2 4 5; do run ${wl}_bwd$c $wl --option flat_bwd_cfg=$c; done
  for c in 1 2 4 5; do run ${wl}_fwd$c $wl --option flat_fwd_cfg=$c; done
done
tail -3 $OUT/err.log
