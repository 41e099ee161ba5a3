"""How much on-SM aggregation of grad_value contributions is there to be had on the benchmark's
encoder workload?  For groups of 1 / 4 / 8 / 32 consecutive queries (same head, same level) count the
distinct destination pixels among all bilinear corner references of the group.  Runs on the CPU.

    python tools/dup_rate.py     # unique/total: 0.83 (one query), 0.50 (the 4 queries of a warp),
                                 #               0.38 (8), 0.28 (the 32 queries of a block)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

p = bench.make_problem('encoder_cfg2', seed=1, device='cpu', frames=1)
loc, shapes = p['loc'], p['shapes']
B, Q, M, L, P, _ = loc.shape
for group in (1, 4, 8, 32):
    tot = uniq = 0
    Qg = Q // group * group
    for l in range(L):
        H, W = int(shapes[l, 0]), int(shapes[l, 1])
        x = loc[:, :Qg, :, l, :, 0] * W - 0.5
        y = loc[:, :Qg, :, l, :, 1] * H - 0.5
        x0, y0 = torch.floor(x).long(), torch.floor(y).long()
        ids = []
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy = x0 + dx, y0 + dy
                ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H) & (x > -1) & (x < W) & (y > -1) & (y < H)
                ids.append(torch.where(ok, yy * W + xx, torch.full_like(xx, -1)))
        ids = torch.stack(ids, -1)                                                   # (B, Qg, M, P, 4)
        ids = ids.reshape(B, Qg // group, group, M, P * 4).permute(0, 1, 3, 2, 4)
        s, _ = ids.reshape(B, Qg // group, M, group * P * 4).sort(-1)
        first = torch.ones_like(s, dtype=torch.bool)
        first[..., 1:] = s[..., 1:] != s[..., :-1]
        uniq += int((first & (s >= 0)).sum())
        tot += int((s >= 0).sum())
    print('queries per group %2d   unique / total corner rows %.3f' % (group, uniq / tot))
