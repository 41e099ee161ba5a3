"""Experiment: is the small-Q forward / backward limited by DRAM (random 128-byte rows) or by
something on the chip?  Same number of samples (300 queries x 8 heads x ~340 samples), value
footprint varied through the number of frames (22.7 MB per frame), kernels launched back to back
(no host gaps inside the timed region), same tensors every launch (L2-warm when they fit).

    python tools/exp_footprint.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pavenet_b200 import _capi  # noqa: E402
from pavenet_b200.functional import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402


def problem(T, P, Q=300):
    bench.WORKLOADS['_x'] = dict(desc='x', B=1, T=T, Q=Q, P=P, levels=bench.R50_LEVELS, kind='pose')
    return bench.make_problem('_x', seed=5, device='cuda')


def timeit(fn, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for flat in (1, 0):
    _capi.set_option('flat', flat)
    for T, P in ((1, 85), (2, 42), (3, 28), (5, 17), (8, 11)):
        p = problem(T, P)
        gv = torch.zeros_like(p['value'])
        gl, ga = torch.empty_like(p['loc']), torch.empty_like(p['aw'])
        rows = 4 * 300 * 8 * p['dims']['L'] * P
        f = timeit(lambda: ms_deform_attn_forward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], 64))
        bw = timeit(lambda: ms_deform_attn_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'],
                                                     p['grad_out'], gv, gl, ga, 64))
        z = timeit(lambda: gv.zero_())
        print('flat=%d T=%d P=%2d value %6.1f MB rows %.2fM  fwd %6.1f us (%.2f clk/row/SM)  bwd %6.1f us (%.2f)  zero %5.1f us'
              % (flat, T, P, p['value'].numel() * 4 / 1e6, rows / 1e6, f, f * 1e-6 * 1.965e9 * 148 / rows,
                 bw, bw * 1e-6 * 1.965e9 * 148 / rows, z))
# large-Q reference point: encoder, one frame
p = bench.make_problem('encoder_cfg2', seed=5, device='cuda', frames=1)
f = timeit(lambda: ms_deform_attn_forward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], 64))
print('encoder 1 frame fwd %.1f us' % f)
