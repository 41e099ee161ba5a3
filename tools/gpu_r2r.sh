#!/bin/bash
OUT=gpurun_out/r2r
mkdir -p $OUT
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-28s fwd %.4f zero %.4f bwd %.4f | step %.4f ms (eager %.4f) %s | q/s %.4g frac step %.3f launches %d' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['ms_per_step_eager'], d['launch'][:10], d['value'], d['roofline_step']['frac'], d['gpu_launches']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  run ${wl} $wl
  run ${wl}_bf16 $wl --value-dtype bf16
  run ${wl}_fused $wl --fused
done
run encoder_cfg2_graph encoder_cfg2 --launch graph
run encoder_cfg2 encoder_cfg2
tail -3 $OUT/err.log
