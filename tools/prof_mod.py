import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pavenet_b200
from torch.profiler import profile, ProfilerActivity
dev='cuda'
R50=[(100,167),(50,84),(25,42),(13,21)]
shapes=torch.tensor(R50,device=dev); sizes=shapes[:,0]*shapes[:,1]; lsi=torch.cat([sizes.new_zeros(1),sizes.cumsum(0)[:-1]]); S=int(sizes.sum())
enc=pavenet_b200.MultiScaleDeformableAttention(dropout=0.0).to(dev)
ref=torch.rand(3,S,4,2,device=dev); q=torch.randn(S,3,256,device=dev,requires_grad=True); qp=torch.randn(S,3,256,device=dev)
def step():
    out=enc(q,query_pos=qp,reference_points=ref,spatial_shapes=shapes,level_start_index=lsi); out.sum().backward()
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
rows=[(e.self_device_time_total/10, e.count//10, e.key) for e in prof.key_averages() if e.self_device_time_total>0]
tot=sum(r[0] for r in rows)
for t,c,k in sorted(rows,reverse=True)[:22]: print('%8.1f us x%-3d %s'%(t,c,k[:110]))
print('total %.1f us'%tot)
