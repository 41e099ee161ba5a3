#!/bin/bash
# closing validation: whole GPU suite, smoke, both bench arms, sanitizer over the zero-fill tests (rows TMA fill is new)
OUT=gpurun_out/r2_final3
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 ) 2>&1 | tee $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
r = json.load(open('$OUT/bench_reference.json')); d = json.load(open('$OUT/bench_default.json'))
print('reference', r['value'], r['ms_per_step'], 'config equal', r['config'] == d['config'])
print('ours', d['value'], d['ms_per_step'], d['kernel_ms'], 'frac', d['roofline']['frac'])
e = d['e2e']
print('e2e', e['mode'], e['value'], e['ms_per_step'], 'other', (e.get('blocking') or e.get('queued'))['ms_per_step'], 'ratio e2e', e['value'] / r['value'])
print('pavenet_step', d['pavenet_step']['value'], d['pavenet_step']['ms_per_step'])
PY
timeout 900 python bench.py --steps 200 --warmup 20 > $OUT/bench_default_200.json 2> $OUT/bench_default_200.err
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 86 --launch-timeout 300 python -m pytest tests/test_gpu_kernel_families.py -m gpu -q -x -p no:cacheprovider -k clear > $OUT/san_$tool.log 2>&1
  echo "exit $?" >> $OUT/san_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " $OUT/san_$tool.log | tail -3
done
