import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import _capi
lib = _capi.load()
torch.manual_seed(0)
dev = 'cuda'
for rows in (128, 1000, 66669):
    x = torch.randn(rows, 256, device=dev)
    w = torch.randn(256, 256, device=dev) * 0.06
    b = torch.randn(256, device=dev)
    y = torch.full((rows, 256), float('nan'), device=dev)
    scratch = torch.empty(2 * 256 * 256, device=dev)
    rc = lib.msda_linear256(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, 0, y.data_ptr(), rows, 0,
                            scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print('rc', rc, lib.msda_last_error() if rc else '')
    ref = (x.double() @ w.double().t() + b.double())
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    ref32 = torch.nn.functional.linear(x, w, b)
    err32 = (ref32.double() - ref).abs().max().item() / ref.abs().max().item()
    print('rows', rows, 'rel err ours %.3e  torch fp32 %.3e' % (err, err32), 'nan count', int(torch.isnan(y).sum()))
    if rows == 1000:
        print(y[:2, :6], ref[:2, :6])
# timing
rows = 66669
x = torch.randn(rows, 256, device=dev); w = torch.randn(256, 256, device=dev) * 0.06; b = torch.randn(256, device=dev)
y = torch.empty(rows, 256, device=dev); scratch = torch.empty(2 * 256 * 256, device=dev)
def ours():
    lib.msda_linear256(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, 0, y.data_ptr(), rows, 0, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
def ref():
    torch.nn.functional.linear(x, w, b)
for fn, name in ((ours, 'tcgen05 3xTF32'), (ref, 'torch fp32 (cuBLAS)')):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, '%.4f ms' % (e0.elapsed_time(e1) / 50))

# ---- weight gradient ----
print('== wgrad')
for rows in (16, 100, 1000, 66669):
    dy = torch.randn(rows, 256, device=dev); x = torch.randn(rows, 256, device=dev)
    dw = torch.full((256, 256), float('nan'), device=dev)
    rc = lib.msda_linear256_wgrad(dy.data_ptr(), x.data_ptr(), None, 0, dw.data_ptr(), rows, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = dy.double().t() @ x.double()
    err = (dw.double() - ref).abs().max().item() / ref.abs().max().item()
    err32 = ((dy.t() @ x).double() - ref).abs().max().item() / ref.abs().max().item()
    print('rc', rc, lib.msda_last_error() if rc else '', 'rows', rows, 'rel err ours %.3e torch fp32 %.3e' % (err, err32), 'nan', int(torch.isnan(dw).sum()))
rows = 66669
dy = torch.randn(rows, 256, device=dev); x = torch.randn(rows, 256, device=dev); dw = torch.empty(256, 256, device=dev)
def ours_w():
    lib.msda_linear256_wgrad(dy.data_ptr(), x.data_ptr(), None, 0, dw.data_ptr(), rows, torch.cuda.current_stream().cuda_stream)
def ref_w():
    dy.t() @ x
for fn, name in ((ours_w, 'wgrad tcgen05 3xTF32'), (ref_w, 'wgrad torch fp32 (cuBLAS)')):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, '%.4f ms' % (e0.elapsed_time(e1) / 50))
