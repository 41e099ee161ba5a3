#!/bin/bash
# backward variant 3 (coarsest level privatised in shared memory, CAS.128): parity, timing, red sectors
OUT=gpurun_out/r3r
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernel_families.py tests/test_gpu_full_size.py -m gpu -q --tb=short -x -k "variant or encoder_cfg2" 2>&1 | tail -3
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-22s fwd %.4f zero %.4f bwd %.4f step %.4f' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
run enc_default encoder_cfg2
run enc_priv encoder_cfg2 --option bwd_variant=3
run enc_priv_off encoder_cfg2 --option bwd_variant=3 --option agg_tile_kb=1
run enc_default2 encoder_cfg2
run enc_priv2 encoder_cfg2 --option bwd_variant=3
run stress_priv stress_cfg5 --option bwd_variant=3
run rand_priv encoder_cfg2_rand --option bwd_variant=3
timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:msda_bwd -s 3 -c 1 --csv --log-file $OUT/ncu_priv.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --option bwd_variant=3 > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('$OUT/ncu_priv.csv')) if len(r) > 10]
for r in rows[1:]:
    print(r[4][:40], r[-3], r[-1])
PY
tail -2 $OUT/err.log
