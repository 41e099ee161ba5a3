#!/bin/bash
# Quick GPU iteration: parity tests + kernel timings for all workloads (no e2e / CPU legs).
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
for wl in encoder_cfg2 pose_cfg3 pose_cfg3_t3 petr_cfg1 stress_cfg5; do
  timeout 300 python bench.py --steps 100 --warmup 10 --workload $wl --no-cpu-baseline --no-e2e 2>>$OUT/bench.err > $OUT/bench_$wl.json
done
timeout 300 python bench.py --steps 100 --warmup 10 --value-dtype bf16 --no-cpu-baseline --no-e2e 2>>$OUT/bench.err > $OUT/bench_bf16.json
timeout 300 python bench.py --steps 100 --warmup 10 --workload pose_cfg3 --value-dtype bf16 --no-cpu-baseline --no-e2e 2>>$OUT/bench.err > $OUT/bench_pose_cfg3_bf16.json
python - <<PY
import json,glob
for f in sorted(glob.glob('$OUT/bench*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f, 'ERR', e); continue
    k=d['kernel_ms']
    print('%-28s q/s %.4g  step %.4f ms  fwd %.4f zero %.4f bwd %.4f  frac bwd %.3f fwd %.3f step %.3f' % (f.split('/')[-1], d['value'], d['ms_per_step'], k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['roofline']['frac'], d['roofline_fwd']['frac'], d['roofline_step']['frac']))
PY
tail -3 $OUT/bench.err
if [ -n "$NCU_FULL" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -f -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ${NCU_ARGS} > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log | cut -c1-200
fi
