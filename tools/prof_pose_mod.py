"""Kernel breakdown of the T-frame pose-decoder attention module, fused prologue on / off."""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pavenet_b200  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 5
K = 17 if T == 5 else 15
dev = 'cuda'
R50 = [(100, 167), (50, 84), (25, 42), (13, 21)]
g = torch.Generator(device=dev).manual_seed(0)
shapes = torch.tensor(R50, device=dev)
sizes = shapes[:, 0] * shapes[:, 1]
lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])
S, C, L = int(sizes.sum()), 256, 4
cls = getattr(pavenet_b200, 'MulFramesMultiScaleDeformablePoseAttentionNumFrames%d' % T)
pose = cls(num_points=K, dropout=0.0).to(dev)
with torch.no_grad():
    for n, p in pose.named_parameters():
        if 'sampling_offsets' in n:
            p.add_(torch.randn_like(p) * 0.02)
pq = torch.randn(300, 1, C, generator=g, device=dev).requires_grad_()
pqp = torch.randn(300, 1, C, generator=g, device=dev)
mem = torch.randn(S, T, C, generator=g, device=dev).requires_grad_()
centre = torch.rand(1, T * 300, 1, 1, 2, generator=g, device=dev) * 0.6 + 0.2
kp = centre + (torch.rand(1, T * 300, 1, K, 2, generator=g, device=dev) - 0.5) * 0.3
rp = kp.expand(1, T * 300, L, K, 2).reshape(1, T * 300, L, 2 * K).contiguous().requires_grad_()


def step():
    out = pose(pq, None, mem, query_pos=pqp, reference_points=rp, spatial_shapes=shapes,
               level_start_index=lsi)
    out.sum().backward()


for fuse in (True, False):
    pose.fuse_prologue = fuse
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step()
    e1.record()
    torch.cuda.synchronize()
    print('== T=%d fuse_prologue=%s  %.3f ms / fwd+bwd' % (T, fuse, e0.elapsed_time(e1) / 30))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    rows = [(e.self_device_time_total / 10, e.count // 10, e.key) for e in prof.key_averages()
            if e.self_device_time_total > 0]
    for t, c, k in sorted(rows, reverse=True)[:14]:
        print('%8.1f us x%-3d %s' % (t, c, k[:120]))
    print('total device %.1f us' % sum(r[0] for r in rows))
