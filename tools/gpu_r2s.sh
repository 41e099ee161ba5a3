#!/bin/bash
OUT=gpurun_out/r2s
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernel_families.py -m gpu -q --tb=short -x 2>&1 | tail -3 | tee $OUT/pytest.log
B="python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-gpu-baseline --model-steps 0"
run() { tag=$1; wl=$2; shift 2
  timeout 300 $B --workload $wl "$@" 2>>$OUT/err.log > $OUT/$tag.json
  python - <<PY
import json
try:
    d = json.load(open('$OUT/$tag.json')); k = d['kernel_ms']
    print('%-28s fwd %.4f zero %.4f bwd %.4f | step %.4f ms (eager %.4f) | frac step %.3f' % ('$tag', k['fwd'], k['grad_value_zero_fill'], k['bwd'], d['ms_per_step'], d['ms_per_step_eager'], d['roofline_step']['frac']))
except Exception as e:
    print('$tag', 'ERR', e)
PY
}
for wl in pose_cfg3 pose_cfg3_t3; do
  for h in 0 1 2 3; do
    run ${wl}_hint$h $wl --option flat_l2_hint=$h
  done
done
# parity with hints on
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from pavenet_b200 import _capi
from pavenet_b200.functional import ms_deform_attn_backward
p = bench.make_problem('pose_cfg3_t3', seed=3, device='cuda')
res = {}
for h in (0, 3):
    _capi.set_option('flat_l2_hint', h)
    gv = torch.zeros_like(p['value']); gl = torch.empty_like(p['loc']); ga = torch.empty_like(p['aw'])
    ms_deform_attn_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], p['grad_out'], gv, gl, ga, 64)
    torch.cuda.synchronize(); res[h] = (gv, gl, ga)
print('hint 3 vs 0', [float((a - b).abs().max() / b.abs().max()) for a, b in zip(res[3], res[0])])
PY
tail -3 $OUT/err.log
