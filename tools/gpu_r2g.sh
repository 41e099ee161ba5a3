#!/bin/bash
OUT=gpurun_out/r2g
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernel_families.py "tests/test_gpu_parity.py::test_clip_model_graphed_step_equals_eager" -m gpu -q --tb=short 2>&1 | tail -15 | tee $OUT/pytest.log
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
run enc_base encoder_cfg2
run enc_fwd3 encoder_cfg2 --option fwd_variant=3
run enc_fwd3_fused encoder_cfg2 --option fwd_variant=3 --fused
run enc_base_fused encoder_cfg2 --fused
run stress_base stress_cfg5
run stress_fwd3 stress_cfg5 --option fwd_variant=3
run pose_cfg3_base pose_cfg3
run pose_cfg3_cfg6 pose_cfg3 --option flat_fwd_cfg=6
run pose_t3_base pose_cfg3_t3
run pose_t3_cfg6 pose_cfg3_t3 --option flat_fwd_cfg=6
run petr_base petr_cfg1
run petr_cfg6 petr_cfg1 --option flat_fwd_cfg=6
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:msda_fwd -s 4 -c 1 --csv --log-file $OUT/ncu_enc_fwd3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --option fwd_variant=3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2g/ncu_enc_fwd3.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d.get('Kernel Name','')[:50], d.get('Metric Name'), d.get('Metric Value'))
PY
tail -3 $OUT/err.log
