#!/bin/bash
# Final validation of a round: whole GPU suite, smoke, the driver's two bench arms, all op workloads.
OUT=gpurun_out/${1:-final}
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 ) 2>&1 | tee $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
r = json.load(open('$OUT/bench_reference.json')); d = json.load(open('$OUT/bench_default.json'))
print('reference', r['value'], r['ms_per_step'], r.get('same_job_as_ours'), 'config equal', r['config'] == d['config'])
print('ours', d['value'], d['ms_per_step'], d['kernel_ms'], 'frac', d['roofline']['frac'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'ratio e2e', d['e2e']['value'] / r['value'])
print('gpu_baseline', d['gpu_baseline']['kernel_ms'], d['gpu_baseline']['speedup'])
print('pavenet_step', d['pavenet_step']['value'], d['pavenet_step']['ms_per_step'], d['pavenet_step']['gpu_launches'])
PY
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --model-steps 0"
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1 stress_cfg5 stress_cfg5_big; do
  timeout 300 $B --workload $wl > $OUT/bench_$wl.json 2>>$OUT/err.log
done
timeout 300 $B --value-dtype bf16 --no-gpu-baseline > $OUT/bench_encoder_cfg2_bf16.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3 --value-dtype bf16 --no-gpu-baseline > $OUT/bench_pose_cfg3_bf16.json 2>>$OUT/err.log
timeout 300 $B --fused --no-gpu-baseline > $OUT/bench_encoder_cfg2_fused.json 2>>$OUT/err.log
timeout 300 $B --workload pose_cfg3 --fused --no-gpu-baseline > $OUT/bench_pose_cfg3_fused.json 2>>$OUT/err.log
python - <<PY
import json, glob
for f in sorted(glob.glob('$OUT/bench_*.json')):
    try:
        d = json.load(open(f)); k = d['kernel_ms']
    except Exception as e:
        print(f, 'ERR', e); continue
    g = d.get('gpu_baseline') or {}
    print('%-36s q/s %.4g step %.4f fwd %.4f zero %.4f bwd %.4f frac fwd %.3f bwd %.3f step %.3f  vs ref kernels %s' % (
        f.split('/')[-1][6:-5], d['value'], d['ms_per_step'], k['fwd'], k['grad_value_zero_fill'], k['bwd'],
        d['roofline_fwd']['frac'], d['roofline']['frac'], d['roofline_step']['frac'],
        {a: round(b, 2) for a, b in (g.get('speedup') or {}).items()}))
PY
tail -3 $OUT/err.log
