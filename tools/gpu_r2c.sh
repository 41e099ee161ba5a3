#!/bin/bash
# Round 2, run C: kernel-variant tests, flat-kernel config sweep, TMA-reduce microbenchmark,
# encoder variants (head-affine forward, warp-aggregated backward).
OUT=gpurun_out/r2c
mkdir -p $OUT
echo "== pytest families"; timeout 600 python -m pytest tests/test_gpu_kernel_families.py -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== microbench tma red"; (cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_tma_red microbench_tma_red.cu && timeout 300 ./microbench_tma_red) 2>&1 | tee $OUT/microbench_tma_red.txt
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
for c in 0 1 2 3 4 5; do
  run pose_cfg3_cfg$c pose_cfg3 --fold-clear 0 --option flat_fwd_cfg=$c --option flat_bwd_cfg=$c
done
for c in 1 2 3 5; do
  run pose_cfg3_t3_cfg$c pose_cfg3_t3 --fold-clear 0 --option flat_fwd_cfg=$c --option flat_bwd_cfg=$c
done
run enc_base encoder_cfg2
run enc_fwd2 encoder_cfg2 --option fwd_variant=2
run enc_bwd2 encoder_cfg2 --option bwd_variant=2
run enc_bwd2_l2 encoder_cfg2 --option bwd_variant=2 --option agg_min_level=2
run enc_bwd2_l3 encoder_cfg2 --option bwd_variant=2 --option agg_min_level=3
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:msda -s 8 -c 2 --csv --log-file $OUT/ncu_enc_base.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_sector_hit_rate.pct,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:msda -s 8 -c 2 --csv --log-file $OUT/ncu_enc_var2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --option fwd_variant=2 --option bwd_variant=2 > /dev/null 2>&1
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2c/ncu_enc_*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    if not rows: print(f,'empty'); continue
    h=rows[0]
    for r in rows[1:]:
        d=dict(zip(h,r)); print(f.split('/')[-1], d.get('Kernel Name','')[:40], d.get('Metric Name'), d.get('Metric Value'))
PY
tail -5 $OUT/err.log
