#!/bin/bash
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernel_families.py tests/test_gpu_full_size.py -m gpu -q --tb=short -k "tile or encoder_cfg2" 2>&1 | tail -25 | tee $OUT/pytest.log
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-34s fwd %.4f  zero %.4f  bwd %.4f  step %.4f ms   frac step %.3f  fwd frac %.3f' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['roofline_step']['frac'], d['roofline_fwd']['frac']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
run enc_base encoder_cfg2
run enc_tile encoder_cfg2 --option fwd_variant=5
run stress_tile stress_cfg5 --option fwd_variant=5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd -s 4 -c 1 -f -o $OUT/prof_tile python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --option fwd_variant=5 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log | cut -c1-200
bash tools/sanitize_r2.sh r2k
tail -3 $OUT/err.log
