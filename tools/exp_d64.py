"""Encoder-shaped problem with 4 heads x 64 channels (embed 256): forward / backward kernel times of the
loaded library (A/B of the D = 64 backward's resident blocks).  Usage: python tools/exp_d64.py"""
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from pavenet_b200.functional import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402

bench.M_HEADS, bench.D_HEAD = 4, 64
dev = torch.device('cuda', 0)
for wl in ('encoder_cfg2', 'pose_cfg3'):
    probs = [bench.make_problem(wl, seed=i, device=dev) for i in range(4)]
    gv = [torch.empty_like(p['value']) for p in probs]
    gl = [torch.empty_like(p['loc']) for p in probs]
    ga = [torch.empty_like(p['aw']) for p in probs]

    def step(i, ev=None):
        p = probs[i % 4]
        if ev: ev[0].record()
        ms_deform_attn_forward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], 64)
        if ev: ev[1].record()
        gv[i % 4].zero_()
        if ev: ev[2].record()
        ms_deform_attn_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], p['grad_out'], gv[i % 4], gl[i % 4], ga[i % 4], 64)
        if ev: ev[3].record()
    for i in range(10):
        step(i)
    torch.cuda.synchronize()
    n = 100
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
    for i in range(n):
        step(i, evs[i])
    torch.cuda.synchronize()
    f = sum(e[0].elapsed_time(e[1]) for e in evs) / n
    b = sum(e[2].elapsed_time(e[3]) for e in evs) / n
    print('%-14s M=4 D=64: fwd %.4f ms  bwd %.4f ms' % (wl, f, b), flush=True)
