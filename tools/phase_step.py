"""Where a PAVE-Net training step spends its time: per phase, host time to *issue* the work
(no synchronisation) and time until the GPU has *finished* it (synchronised), so launch-bound
phases (issue ~= finish) stand out from GPU-bound ones.  Run on a B200."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import clip_model  # noqa: E402

torch.manual_seed(0)
dev = torch.device('cuda:0')
model = clip_model.PaveNetR50().to(dev).train()
if len(sys.argv) > 1 and sys.argv[1] == 'graphs':
    model.enable_graphs()
opt = clip_model.build_optimizer(model)
batch = clip_model.synthetic_clip_batch(1, dev, seed=1)
params = [p for p in model.parameters() if p.requires_grad]


def run(sync):
    marks = []

    def hook(name):
        if sync:
            torch.cuda.synchronize()
        marks.append((name, time.perf_counter()))

    model.phase_hook = hook
    if model._graphed is not None:
        from pavenet_b200 import graphs
        graphs.refresh_seed(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = model(*batch)
    loss = sum(losses.values())
    opt.zero_grad(set_to_none=True)
    hook('loss sum')
    loss.backward()
    hook('backward')
    torch.nn.utils.clip_grad_norm_(params, 0.1)
    opt.step()
    hook('clip + AdamW')
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    model.phase_hook = None
    out, prev = [], t0
    for name, t in marks:
        out.append((name, (t - prev) * 1e3))
        prev = t
    return out, (t_end - t0) * 1e3


for _ in range(3):
    clip_model.train_step(model, opt, *batch)
res = {}
for sync in (False, True):
    acc, tot = None, 0.0
    for _ in range(5):
        o, t = run(sync)
        tot += t
        acc = o if acc is None else [(n, a + b) for (n, a), (_, b) in zip(acc, o)]
    res[sync] = ([(n, a / 5) for n, a in acc], tot / 5)
print('%-28s %12s %14s' % ('phase', 'issue ms', 'finished ms'))
for (n, a), (_, b) in zip(res[False][0], res[True][0]):
    print('%-28s %12.2f %14.2f' % (n, a, b))
print('%-28s %12.2f %14.2f   (step wall: unsynchronised / synchronised at every phase)' %
      ('total', res[False][1], res[True][1]))
