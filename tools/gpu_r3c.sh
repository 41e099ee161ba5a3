#!/bin/bash
# folded zero-fill left in L2 (evict-last) x L2 hints in the flat backward
OUT=gpurun_out/r3c
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
  run ${wl}_p0h0 $wl
  for h in 0 1 2 3; do run ${wl}_p2h$h $wl --option clear_policy=2 --option flat_l2_hint=$h; done
done
tail -3 $OUT/err.log
