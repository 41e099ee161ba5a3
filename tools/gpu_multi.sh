#!/bin/bash
# multi-GPU validation: N = $1 ranks on one host
N=${1:-2}
OUT=gpurun_out/r2_n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== overlap check"
for m in flat_eager overlap_eager; do timeout 300 $TR tools/check_overlap.py $m > $OUT/check_$m.log 2>&1 || tail -5 $OUT/check_$m.log; done
python tools/check_overlap.py compare 2>&1 | tee $OUT/check_overlap.txt
echo "== default bench line"; ( time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | tail -4
python - <<PY
import json
d = json.load(open('$OUT/bench.json'))
print('op value', d['value'], 'ms', d['ms_per_step'], 'n', d['n_gpus'])
print('e2e', {k: d['e2e'].get(k) for k in ('value', 'ms_per_step', 'host_affinity')})
ps = d['pavenet_step']
print('pavenet_step', {k: ps.get(k) for k in ('value', 'ms_per_step', 'collective', 'gpu_launches', 'error')})
PY
tail -3 $OUT/bench.err
for ge in overlap flat; do
  timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 6 --workload pavenet_step --grad-exchange $ge > $OUT/step_$ge.json 2>> $OUT/err.log
  python -c "
import json; d=json.load(open('$OUT/step_$ge.json')); print('$ge', round(d['value'],2), 'clips/s', round(d['ms_per_step'],2), 'ms', d['collective'])"
done
echo "== host link, all ranks at once"
timeout 300 $TR tools/pcie_duplex.py 2>/dev/null | grep "^rank" | sort | tee $OUT/pcie_unbound.txt
timeout 300 $TR tools/pcie_duplex.py --bind 2>/dev/null | grep "^rank" | sort | tee $OUT/pcie_bound.txt
tail -3 $OUT/err.log
