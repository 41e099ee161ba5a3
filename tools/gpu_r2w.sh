#!/bin/bash
OUT=gpurun_out/r2w
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernel_families.py -m gpu -q --tb=short -x -k "clear" 2>&1 | tail -3 | tee $OUT/pytest.log
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
run enc_sep encoder_cfg2
run enc_fold_memset encoder_cfg2 --fold-clear 1
run enc_fold_tma encoder_cfg2 --fold-clear 1 --option clear_mode=2
run enc_sep2 encoder_cfg2
run enc_fold_tma2 encoder_cfg2 --fold-clear 1 --option clear_mode=2
run big_sep stress_cfg5_big
run big_fold_tma stress_cfg5_big --fold-clear 1 --option clear_mode=2
run enc_rand encoder_cfg2_rand
tail -3 $OUT/err.log
