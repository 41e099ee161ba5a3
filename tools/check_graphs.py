"""Graphed vs eager PAVE-Net step: same losses and gradients (dropout off), then timing."""
import os, sys, time, copy
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import clip_model
torch.manual_seed(0)
dev = torch.device('cuda:0')
model = clip_model.PaveNetR50().to(dev).train()
for m in model.modules():                      # deterministic comparison: no dropout anywhere
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if isinstance(m, torch.nn.MultiheadAttention): m.dropout = 0.0
    if hasattr(m, 'ffn_drop'): m.ffn_drop = 0.0
batch = clip_model.synthetic_clip_batch(1, dev, seed=1)
params = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
def run():
    for _, p in params: p.grad = None
    losses = model(*batch)
    loss = sum(losses.values()); loss.backward()
    return float(loss), {n: p.grad.clone() for n, p in params if p.grad is not None}
l0, g0 = run()
model.enable_graphs()
l1, g1 = run()          # captures
l2, g2 = run()          # replays
print('loss eager %.6f  graphed(capture step) %.6f  graphed(replay) %.6f' % (l0, l1, l2))
worst = 0.0
for n in g0:
    a, b = g0[n], g2[n]
    e = (a - b).abs().max().item() / (a.abs().max().item() + 1e-12)
    worst = max(worst, e)
    if e > 1e-3: print('  grad mismatch', n, e)
print('missing grads in graphed:', [n for n in g0 if n not in g2][:5], 'worst rel grad err %.2e' % worst)
