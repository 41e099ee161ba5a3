#!/bin/bash
# host pipeline: a piece's copies as one cudaMemcpyBatchAsync instead of three cudaMemcpyAsync
OUT=gpurun_out/r3b
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "host_buffer" 2>&1 | tail -2
PAVENET_MSDA_BATCH_COPIES=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "host_buffer" 2>&1 | tail -2
for bc in 0 1 0 1; do
PAVENET_MSDA_BATCH_COPIES=$bc timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline --model-steps 0 > $OUT/bench_$bc.json 2> $OUT/bench_$bc.err
python - <<PY
import json
d = json.load(open('$OUT/bench_$bc.json')); e = d['e2e']
o = e.get('blocking') or e.get('queued')
print('batch copies $bc: e2e %s %.3f ms/step   other form %.3f ms/step   diff %g' % (e['mode'], e['ms_per_step'], o['ms_per_step'], e['max_abs_diff_vs_blocking']))
PY
done
