import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import functional as Fn
torch.manual_seed(0)
rows = 66669
x = torch.randn(rows, 256, device='cuda'); go = torch.randn(rows, 256, device='cuda')
w1 = torch.randn(1024, 256, device='cuda') * 0.06; b1 = torch.randn(1024, device='cuda') * 0.1
w2 = torch.randn(256, 1024, device='cuda') * 0.03; b2 = torch.randn(256, device='cuda') * 0.1
def rel(a, b): return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
h = Fn._linear_fused_raw(x, w1, b1, relu=True)
h_ref = torch.relu(x @ w1.t() + b1)
print('h', rel(h, h_ref))
y = Fn._linear_fused_raw(h, w2, b2, residual=x)
print('y', rel(y, x + h_ref @ w2.t() + b2))
dz = Fn._linear_fused_raw(go, w2.t().contiguous(), gate=h, gate_scale=1.0)
dz_ref = (go @ w2) * (h_ref > 0)
print('dz', rel(dz, dz_ref), 'mismatch gate', int(((h > 0) != (h_ref > 0)).sum()))
for trial in range(3):
    gx = Fn._linear_fused_raw(dz, w1.t().contiguous(), residual=go)
    gx_ref = go + dz @ w1
    d = (gx - gx_ref).abs()
    print('gx', rel(gx, gx_ref), 'bad elems', int((d > 1e-3).sum()), 'rows', torch.unique((d > 1e-3).nonzero()[:, 0])[:10].tolist(),
          'cols', torch.unique((d > 1e-3).nonzero()[:, 1])[:10].tolist())
    gx2 = Fn._linear_fused_raw(dz, w1.t().contiguous())
    print('gx no residual', rel(gx2, dz @ w1))
dw1 = Fn._wgrad_raw(dz, x, None, 0, 1024, 256); print('dw1', rel(dw1, dz.t() @ x))
dw2 = Fn._wgrad_raw(go, h, None, 0, 256, 1024); print('dw2', rel(dw2, go.t() @ h))
print('db1', rel(Fn._colsum_raw(dz), dz.sum(0)))
