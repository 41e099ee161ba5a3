#!/bin/bash
# Round 2, run B: rest of the GPU suite, small-Q with L2 prefetch / warm L2, ncu on the flat kernels.
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e"
run() { # tag workload extra...
  tag=$1; wl=$2; shift 2
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
  run ${wl}_sep $wl --fold-clear 0
  run ${wl}_sep_pf1 $wl --fold-clear 0 --option l2_prefetch=1
  run ${wl}_sep_pf3 $wl --fold-clear 0 --option l2_prefetch=3
  run ${wl}_fold_pf3 $wl --option l2_prefetch=3
  run ${wl}_sep_warm $wl --fold-clear 0 --sets 1
  run ${wl}_fold_warm $wl --sets 1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda -s 8 -c 2 -f -o $OUT/prof_pose python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload pose_cfg3 --fold-clear 0 > $OUT/ncu_pose.log 2>&1
tail -2 $OUT/ncu_pose.log | cut -c1-200
tail -5 $OUT/err.log
