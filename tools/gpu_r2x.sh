#!/bin/bash
OUT=gpurun_out/r2x
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "host_buffer" 2>&1 | tail -5 | tee $OUT/pytest.log
for pm in 0 24; do
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-gpu-baseline --model-steps 0 --piece-mb $pm > $OUT/bench_$pm.json 2> $OUT/bench_$pm.err
python - <<PY
import json
d = json.load(open('$OUT/bench_$pm.json')); e = d['e2e']
print('piece $pm: e2e queued %.3f ms/step (%.4g q/s)   blocking %.3f ms/step   autograd %.3f   diff %g' % (e['ms_per_step'], e['value'], e['blocking']['ms_per_step'], d['e2e_autograd']['ms_per_step'], e['max_abs_diff_vs_blocking']))
PY
tail -2 $OUT/bench_$pm.err
done
