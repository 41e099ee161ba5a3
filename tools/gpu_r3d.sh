#!/bin/bash
OUT=gpurun_out/r3d
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 ) | tee $OUT/pytest.log
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 300 python -m pytest tests/test_gpu_kernel_families.py -m gpu -q -x -p no:cacheprovider -k clear > $OUT/san_$tool.log 2>&1
  echo "exit $?" >> $OUT/san_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " $OUT/san_$tool.log | tail -3
done
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --model-steps 0"
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  timeout 300 $B --workload $wl > $OUT/bench_$wl.json 2>>$OUT/err.log
done
timeout 300 $B --workload pose_cfg3 --value-dtype bf16 --no-gpu-baseline > $OUT/bench_pose_cfg3_bf16.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3_t3 --value-dtype bf16 --no-gpu-baseline > $OUT/bench_pose_cfg3_t3_bf16.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3 --fused --no-gpu-baseline > $OUT/bench_pose_cfg3_fused.json 2>>$OUT/err.log
timeout 300 $B --workload petr_cfg1 --fused --no-gpu-baseline > $OUT/bench_petr_cfg1_fused.json 2>>$OUT/err.log
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json, glob
for f in sorted(glob.glob('$OUT/bench_*.json')):
    try:
        d = json.load(open(f)); k = d['kernel_ms']
    except Exception as e:
        print(f, 'ERR', e); continue
    e = d.get('e2e') or {}
    print('%-28s q/s %.4g step %.4f (eager %.4f) fwd %.4f zero %.4f bwd %.4f frac step %.3f  e2e %s %s' % (
        f.split('/')[-1][6:-5], d['value'], d['ms_per_step'], d['ms_per_step_eager'], k['fwd'], k['grad_value_zero_fill'], k['bwd'],
        d['roofline_step']['frac'], e.get('mode'), e.get('ms_per_step')))
PY
tail -3 $OUT/err.log
