#!/bin/bash
# N = 4 point of the scaling record: the driver's default bench line under torchrun + the clip step alone + a pose workload (graph leg under N ranks)
N=${1:-4}
OUT=gpurun_out/r2_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== default bench line"; ( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | tail -4
python - <<PY
import json
d = json.load(open('$OUT/bench.json'))
print('op value', d['value'], 'ms', d['ms_per_step'], 'n', d['n_gpus'])
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step')})
ps = d['pavenet_step']
print('pavenet_step', {k: ps.get(k) for k in ('value', 'ms_per_step', 'collective', 'gpu_launches', 'error')})
PY
tail -3 $OUT/bench.err
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 6 --workload pavenet_step --grad-exchange flat > $OUT/step_flat.json 2>> $OUT/err.log
python -c "
import json; d=json.load(open('$OUT/step_flat.json')); print('flat', round(d['value'],2), 'clips/s', round(d['ms_per_step'],2), 'ms', d['collective'])"
timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 20 --workload pose_cfg3 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0 > $OUT/pose.json 2>> $OUT/err.log
python -c "
import json; d=json.load(open('$OUT/pose.json')); print('pose_cfg3 x$N', d['value'], 'q/s', d['ms_per_step'], 'ms', d['launch'])"
tail -3 $OUT/err.log
