import sys, os, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import clip_model
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0)
dev = torch.device('cuda:0')
model = clip_model.PaveNetR50().to(dev).train()
opt = clip_model.build_optimizer(model)
batch = clip_model.synthetic_clip_batch(1, dev, seed=1)
for _ in range(3):
    clip_model.train_step(model, opt, *batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        clip_model.train_step(model, opt, *batch)
    torch.cuda.synchronize()
cat = collections.Counter(); cnt = collections.Counter()
def classify(n):
    if 'msda_fwd' in n or 'msda_bwd' in n: return 'msda sampling kernels (ours)'
    if 'linear256' in n or 'split_weight' in n: return 'linear256 tcgen05 (ours)'
    if 'sgemm' in n or 'gemm' in n.lower() or 'cutlass' in n.lower() or 'gemv' in n.lower(): return 'cuBLAS GEMM (fp32)'
    if 'cudnn' in n.lower() or 'conv' in n.lower() or 'xmma' in n.lower() or 'implicit' in n.lower(): return 'cuDNN conv'
    if 'elementwise' in n or 'vectorized' in n or 'reduce_kernel' in n or 'layer_norm' in n.lower() or 'softmax' in n.lower() or 'cat' in n.lower() or 'copy' in n.lower() or 'Memcpy' in n or 'Memset' in n or 'index' in n.lower() or 'scatter' in n.lower() or 'gather' in n.lower() or 'sort' in n.lower() or 'fused_adam' in n.lower() or 'multi_tensor' in n.lower() or 'group_norm' in n.lower() or 'RowwiseMoments' in n or 'batch_norm' in n.lower() or 'dropout' in n.lower() or 'nll' in n.lower(): return 'elementwise / norm / copy / optimizer'
    if 'attention' in n.lower() or 'fmha' in n.lower() or 'flash' in n.lower(): return 'MHA kernels'
    return 'other'
tot = 0
for e in prof.key_averages():
    t = e.self_device_time_total
    if t <= 0: continue
    c = classify(e.key); cat[c] += t; cnt[c] += e.count; tot += t
for c, t in cat.most_common():
    print('%-42s %8.2f ms/step  %5.1f%%  (%d launches/step)' % (c, t / 3e3, 100 * t / tot, cnt[c] // 3))
print('total device time %.2f ms/step' % (tot / 3e3))
others = [(e.self_device_time_total, e.key) for e in prof.key_averages() if classify(e.key) == 'other']
for t, k in sorted(others, reverse=True)[:8]: print('   other:', round(t / 3e3, 3), k[:90])
print('-- top kernels by device time per step')
rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0),
              reverse=True)
for t, c, k in rows[:32]:
    print('  %8.3f ms  x%-5d %s' % (t / 3e3, c // 3, k[:110]))
