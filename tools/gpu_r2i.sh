#!/bin/bash
# 2 GPUs: overlapped gradient exchange check, the driver's N=2 bench invocation, encoder fwd variant 4 (rank 0 only)
OUT=gpurun_out/r2i
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/check_overlap.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -8 | tee $OUT/check_overlap.txt
( time timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err ) 2>&1 | tail -4
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2i/bench_n2.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'n', d['n_gpus'])
print('e2e', d['e2e'])
print('pavenet_step', {k: d['pavenet_step'].get(k) for k in ('value', 'ms_per_step', 'collective', 'gpu_launches', 'error')})
PY
tail -5 $OUT/bench_n2.err
for ge in overlap flat; do
  timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 6 --workload pavenet_step --grad-exchange $ge > $OUT/step_n2_$ge.json 2>> $OUT/err.log
  python -c "
import json; d=json.load(open('$OUT/step_n2_$ge.json')); print('$ge', d['value'], 'clips/s', d['ms_per_step'], 'ms', d['collective'])"
done
timeout 300 python bench.py --steps 20 --warmup 6 --workload pavenet_step > $OUT/step_n1.json 2>> $OUT/err.log; python -c "
import json; d=json.load(open('$OUT/step_n1.json')); print('n1', d['value'], 'clips/s', d['ms_per_step'], 'ms', d['gpu_launches'])"
for v in 0 4; do
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 --option fwd_variant=$v > $OUT/enc_v$v.json 2>>$OUT/err.log; python -c "
import json; d=json.load(open('$OUT/enc_v$v.json')); print('fwd_variant=$v', d['kernel_ms'])"
done
tail -3 $OUT/err.log
