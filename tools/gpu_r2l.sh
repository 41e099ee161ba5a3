#!/bin/bash
OUT=gpurun_out/r2l
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
run enc_tile5 encoder_cfg2 --option fwd_variant=5
run enc_tile6 encoder_cfg2 --option fwd_variant=6
run stress_tile6 stress_cfg5 --option fwd_variant=6
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio --clock-control none -k regex:msda_fwd -s 4 -c 1 --csv --log-file $OUT/ncu_tile6.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --option fwd_variant=6 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2l/ncu_tile6.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d.get('Kernel Name','')[:50], d.get('Metric Name'), d.get('Metric Value'))
PY
tail -3 $OUT/err.log
