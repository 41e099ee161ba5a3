#!/bin/bash
OUT=gpurun_out/r2y
mkdir -p $OUT
for pm in 16 32 48 64 96; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline --model-steps 0 --piece-mb $pm > $OUT/bench_$pm.json 2> $OUT/bench_$pm.err
python - <<PY
import json
d = json.load(open('$OUT/bench_$pm.json')); e = d['e2e']
print('piece $pm MiB: e2e queued %.3f ms/step (%.4g q/s)   blocking %.3f ms/step' % (e['ms_per_step'], e['value'], e['blocking']['ms_per_step']))
PY
done
