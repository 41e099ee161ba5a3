#!/usr/bin/env python
"""Module-level timing: the attention module classes (projections + op +
output projection), forward + backward, with the fused prologue on and off.

    python tools/bench_modules.py            # prints one JSON line per case
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pavenet_b200  # noqa: E402

R50 = [(100, 167), (50, 84), (25, 42), (13, 21)]


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    shapes = torch.tensor(R50, device=dev)
    sizes = shapes[:, 0] * shapes[:, 1]
    lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])
    S, C, L = int(sizes.sum()), 256, 4

    def rnd(*s):
        return torch.randn(*s, generator=g, device=dev)

    cases = {}
    # spatial encoder layer attention: 3 frames, queries = pixels
    enc = pavenet_b200.MultiScaleDeformableAttention(dropout=0.0).to(dev)
    refs = []
    for h, w in R50:
        ys = (torch.arange(h, device=dev) + 0.5) / h
        xs = (torch.arange(w, device=dev) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        refs.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    ref = torch.cat(refs)[None, :, None, :].expand(3, S, L, 2).contiguous()
    q, qp = rnd(S, 3, C).requires_grad_(), rnd(S, 3, C)
    cases['encoder_layer_attention B=3 Q=S=22223'] = (
        enc, lambda m: m(q, query_pos=qp, reference_points=ref, spatial_shapes=shapes,
                         level_start_index=lsi))
    # pose decoder layer attention, T = 5 frames, 300 pose queries x 17 keypoints
    for T, K in ((5, 17), (3, 15)):
        cls = getattr(pavenet_b200, 'MulFramesMultiScaleDeformablePoseAttentionNumFrames%d' % T)
        pose = cls(num_points=K, dropout=0.0).to(dev)
        with torch.no_grad():
            for n, p in pose.named_parameters():
                if 'sampling_offsets' in n:
                    p.add_(torch.randn_like(p) * 0.02)
        pq, pqp = rnd(300, 1, C).requires_grad_(), rnd(300, 1, C)
        mem = rnd(S, T, C).requires_grad_()
        centre = torch.rand(1, T * 300, 1, 1, 2, generator=g, device=dev) * 0.6 + 0.2
        kp = (centre + (torch.rand(1, T * 300, 1, K, 2, generator=g, device=dev) - 0.5) * 0.3)
        rp = kp.expand(1, T * 300, L, K, 2).reshape(1, T * 300, L, 2 * K).contiguous().requires_grad_()
        cases['pose_decoder_attention T=%d Q=300 K=%d' % (T, K)] = (
            pose, lambda m, pq=pq, pqp=pqp, mem=mem, rp=rp: m(
                pq, None, mem, query_pos=pqp, reference_points=rp, spatial_shapes=shapes,
                level_start_index=lsi))

    for name, (mod, call) in cases.items():
        row = {'case': name}
        for fuse in (False, True):
            mod.fuse_prologue = fuse

            def step():
                out = call(mod)
                out.sum().backward()
            row['fused_ms' if fuse else 'op_by_op_ms'] = round(timed(step), 4)
        if hasattr(mod, 'fused'):
            mod.fuse_prologue = False
            mod.fused = False            # the reference's T separate op calls + Z-weighted fusion

            def step():
                out = call(mod)
                out.sum().backward()
            row['reference_style_T_calls_ms'] = round(timed(step), 4)
            mod.fused = True
        row['speedup_fused_vs_op_by_op'] = round(row['op_by_op_ms'] / row['fused_ms'], 3)
        print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
