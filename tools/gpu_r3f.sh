#!/bin/bash
OUT=gpurun_out/r3f
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "host_buffer" 2>&1 | tail -2
for cfg in "2 0" "2 4096" "3 4096" "4 4096" "3 0"; do
set -- $cfg
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline --model-steps 0 --e2e-depth $1 --piece-mb $2 > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err
python - <<PY
import json
d = json.load(open('$OUT/bench_$1_$2.json')); e = d['e2e']
q = e if e['mode'] == 'queued' else e['queued']; b = e['blocking'] if e['mode'] == 'queued' else e
print('depth $1 piece $2 MiB: queued %.3f ms/step   blocking %.3f ms/step  diff %g' % (q['ms_per_step'], b['ms_per_step'], e['max_abs_diff_vs_blocking']))
PY
done
