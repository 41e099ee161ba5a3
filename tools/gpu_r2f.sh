#!/bin/bash
# Round 2, run F: the whole GPU suite, smoke, and the driver's default bench invocations.
OUT=gpurun_out/r2f
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 ) 2>&1 | tee $OUT/pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -5 | tee $OUT/smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2f/bench_default.json'))
for k in ('value', 'ms_per_step', 'kernel_ms', 'gpu_launches', 'clocks'):
    print(k, d.get(k))
print('roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'traffic')})
print('e2e', d['e2e'])
print('cpu_baseline', d.get('cpu_baseline'))
print('gpu_baseline', d.get('gpu_baseline'))
print('pavenet_step', d.get('pavenet_step'))
print('onchip', d.get('roofline_onchip'))
PY
tail -5 $OUT/bench_default.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err ) 2>&1 | tail -4
cat $OUT/bench_reference.json | cut -c1-600
