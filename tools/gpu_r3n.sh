#!/bin/bash
# closing validation (kernels changed again: constant-bank base pointers): suite, smoke, sanitizer, final lines
OUT=gpurun_out/r2_final4
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 ) | tee $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $OUT/smoke.log
bash tools/sanitize_r2.sh r2_final4/san 2>&1 | tail -8
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --model-steps 0"
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1 stress_cfg5 stress_cfg5_big encoder_cfg2_rand; do
  timeout 300 $B --workload $wl > $OUT/bench_$wl.json 2>>$OUT/err.log
done
timeout 300 $B --value-dtype bf16 --no-gpu-baseline > $OUT/bench_encoder_cfg2_bf16.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3 --value-dtype bf16 --no-gpu-baseline > $OUT/bench_pose_cfg3_bf16.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3_t3 --value-dtype bf16 --no-gpu-baseline > $OUT/bench_pose_cfg3_t3_bf16.json 2>>$OUT/err.log
timeout 300 $B --fused --no-gpu-baseline > $OUT/bench_encoder_cfg2_fused.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3 --fused --no-gpu-baseline > $OUT/bench_pose_cfg3_fused.json 2>>$OUT/err.log
timeout 300 $B --workload petr_cfg1 --fused --no-gpu-baseline > $OUT/bench_petr_cfg1_fused.json 2>>$OUT/err.log
python - <<PY
import json, glob
r = json.load(open('$OUT/bench_reference.json'))
print('reference arm', r['value'], r['ms_per_step'])
for f in sorted(glob.glob('$OUT/bench_*.json')):
    if 'reference' in f: continue
    try:
        d = json.load(open(f)); k = d['kernel_ms']
    except Exception as e:
        print(f, 'ERR', e); continue
    e = d.get('e2e') or {}
    g = (d.get('gpu_baseline') or {}).get('speedup') or {}
    print('%-24s q/s %.4g step %.4f (eager %.4f) fwd %.4f zero %.4f bwd %.4f frac f/b/s %.3f %.3f %.3f e2e %s %s vs-ref-kernels %s' % (
        f.split('/')[-1][6:-5], d['value'], d['ms_per_step'], d['ms_per_step_eager'], k['fwd'], k['grad_value_zero_fill'], k['bwd'],
        d['roofline_fwd']['frac'], d['roofline']['frac'], d['roofline_step']['frac'], e.get('mode'), e.get('ms_per_step'),
        {a: round(b, 2) for a, b in g.items()}))
PY
tail -3 $OUT/err.log
