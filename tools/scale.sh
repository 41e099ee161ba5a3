#!/bin/bash
# 1/2/4/8-GPU weak-scaling runs of the op bench and of the PAVE-Net clip step.
OUT=gpurun_out/${1:-scale}
mkdir -p $OUT
for n in 1 2 4 8; do
  for wl in encoder_cfg2 pavenet_step; do
    steps=100; warm=10; [ $wl = pavenet_step ] && steps=15 && warm=4
    if [ $n = 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps $steps --warmup $warm --workload $wl --no-cpu-baseline --no-e2e > $OUT/${wl}_n$n.json 2>$OUT/${wl}_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps $steps --warmup $warm --workload $wl --no-cpu-baseline --no-e2e > $OUT/${wl}_n$n.json 2>$OUT/${wl}_n$n.err
    fi
    python -c "
import json,sys
try:
    d=json.load(open('$OUT/${wl}_n$n.json')); print('$wl N=$n', round(d['value'],3), d['unit'], 'ms/step', round(d['ms_per_step'],4))
except Exception as e: print('$wl N=$n FAILED', e)
"
  done
done
