"""Top ATen / custom ops of a PAVE-Net training step by GPU time, grouped by input shape
(finds the large copies, reductions and GEMMs worth replacing).  Run on a B200."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import clip_model  # noqa: E402

torch.manual_seed(0)
dev = torch.device('cuda:0')
model = clip_model.PaveNetR50().to(dev).train()
opt = clip_model.build_optimizer(model)
batch = clip_model.synthetic_clip_batch(1, dev, seed=1)
for _ in range(3):
    clip_model.train_step(model, opt, *batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    clip_model.train_step(model, opt, *batch)
    torch.cuda.synchronize()
rows = [(e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:120])
        for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print('total self device time %.2f ms' % (tot / 1e3))
for t, c, k, shp in rows[:int(sys.argv[1]) if len(sys.argv) > 1 else 70]:
    print('%8.3f ms x%-4d %-38s %s' % (t / 1e3, c, k[:38], shp))
