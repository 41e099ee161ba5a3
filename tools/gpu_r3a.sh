#!/bin/bash
# does the backward get faster when the zero lines of grad_value are left in L2 by the folded fill?
OUT=gpurun_out/r3a
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
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  for pol in 0 1 2; do run ${wl}_pol$pol $wl --option clear_policy=$pol; done
done
tail -3 $OUT/err.log
