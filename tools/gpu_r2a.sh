#!/bin/bash
# Round 2, run A: whole GPU suite + small-Q A/B (flat kernels on/off, clear folded or separate).
OUT=gpurun_out/r2a
mkdir -p $OUT
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e"
run() { # tag workload extra...
  tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-34s fwd %.4f  zero %.4f  bwd %.4f  step %.4f ms   frac step %.3f  fam %s' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['roofline_step']['frac'], d.get('kernel_families')))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  run ${wl}_rows_sepclear $wl --option flat=0 --fold-clear 0
  run ${wl}_flat_sepclear $wl --fold-clear 0
  run ${wl}_flat_fold $wl
  run ${wl}_flat_fold_bf16 $wl --value-dtype bf16
  run ${wl}_fused_flat_fold $wl --fused
  run ${wl}_fused_rows_sepclear $wl --fused --option flat=0 --fold-clear 0
done
run encoder_cfg2 encoder_cfg2
run encoder_cfg2_fused encoder_cfg2 --fused
tail -5 $OUT/err.log
