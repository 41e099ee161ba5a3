#!/bin/bash
OUT=gpurun_out/r3g
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "host_buffer" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
python - <<PY
import json
d = json.load(open('$OUT/bench_default.json')); e = d['e2e']
o = e.get('blocking') or e.get('queued')
print('default: value %.4g  step %.4f  e2e %s %.3f ms (%.4g q/s)  other %.3f  autograd %.3f  diff %g' % (d['value'], d['ms_per_step'], e['mode'], e['ms_per_step'], e['value'], o['ms_per_step'], d['e2e_autograd']['ms_per_step'], e['max_abs_diff_vs_blocking']))
print(e['api'])
PY
tail -2 $OUT/bench_default.err
