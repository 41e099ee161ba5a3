#!/bin/bash
OUT=gpurun_out/r2o
mkdir -p $OUT
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-34s fwd %.4f  zero %.4f  bwd %.4f  step %.4f ms   frac step %.3f' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['roofline_step']['frac']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
for wl in pose_cfg3 pose_cfg3_t3 petr_cfg1; do
  run ${wl}_base $wl
  run ${wl}_pipe6 $wl --option flat_bwd_cfg=6
  run ${wl}_pipe7 $wl --option flat_bwd_cfg=7
done
# parity of the pipelined variants on one full-size problem
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from pavenet_b200 import _capi
from pavenet_b200.functional import ms_deform_attn_backward
p = bench.make_problem('pose_cfg3_t3', seed=3, device='cuda')
res = {}
for cfg in (0, 6, 7):
    _capi.set_option('flat_bwd_cfg', cfg)
    gv = torch.zeros_like(p['value']); gl = torch.empty_like(p['loc']); ga = torch.empty_like(p['aw'])
    ms_deform_attn_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], p['grad_out'], gv, gl, ga, 64)
    torch.cuda.synchronize(); res[cfg] = (gv, gl, ga)
for cfg in (6, 7):
    print('cfg', cfg, [float((a - b).abs().max() / b.abs().max()) for a, b in zip(res[cfg], res[0])])
PY
# config 5 at 1200x2000: batch sweep
for f in 1 2 4 8 16; do
  timeout 300 $B --workload stress_cfg5_big --frames $f 2>>$OUT/err.log > $OUT/big_f$f.json
  python -c "
import json; d=json.load(open('$OUT/big_f$f.json')); k=d['kernel_ms']; print('stress_cfg5_big frames=$f  q/s %.4g  step %.4f ms  fwd %.4f bwd %.4f  frac step %.3f' % (d['value'], d['ms_per_step'], k['fwd'], k['bwd'], d['roofline_step']['frac']))"
done
tail -3 $OUT/err.log
