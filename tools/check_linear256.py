"""Accuracy and timing of the tcgen05 projections (forward, weight gradient, bias gradient)
for every supported (in, out) against float64 / cuBLAS fp32.  Run on a B200."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pavenet_b200 import _capi  # noqa: E402

lib = _capi.load()
torch.manual_seed(0)
dev = 'cuda'
SHAPES = ((256, 256), (256, 128), (128, 256), (256, 1024), (1024, 256))     # (in, out)


def stream():
    return torch.cuda.current_stream().cuda_stream


def time_ms(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rel(a, ref):
    return (a.double() - ref).abs().max().item() / ref.abs().max().item()


for n_in, n_out in SHAPES:
    print('== forward in=%d out=%d' % (n_in, n_out))
    for rows in (1, 128, 1000, 66669):
        x = torch.randn(rows, n_in, device=dev)
        w = torch.randn(n_out, n_in, device=dev) * 0.06
        b = torch.randn(n_out, device=dev)
        y = torch.full((rows, n_out), float('nan'), device=dev)
        scratch = torch.empty(2 * n_in * n_out, device=dev)
        rc = lib.msda_linear256(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, 0, y.data_ptr(), rows,
                                n_in, n_out, 0, scratch.data_ptr(), stream())
        torch.cuda.synchronize()
        ref = x.double() @ w.double().t() + b.double()
        print('  rc', rc, lib.msda_last_error() if rc else '', 'rows', rows,
              'rel err ours %.3e  torch fp32 %.3e' % (rel(y, ref), rel(torch.nn.functional.linear(x, w, b), ref)),
              'nan', int(torch.isnan(y).sum()))
    t_ours = time_ms(lambda: lib.msda_linear256(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, 0, y.data_ptr(),
                                                rows, n_in, n_out, 0, scratch.data_ptr(), stream()))
    t_ref = time_ms(lambda: torch.nn.functional.linear(x, w, b))
    print('  rows %d: tcgen05 3xTF32 %.4f ms   torch fp32 (cuBLAS) %.4f ms' % (rows, t_ours, t_ref))

    print('== wgrad  in=%d out=%d' % (n_in, n_out))
    for rows in (1, 16, 100, 1000, 66669):
        dy = torch.randn(rows, n_out, device=dev)
        x = torch.randn(rows, n_in, device=dev)
        dw = torch.full((n_out, n_in), float('nan'), device=dev)
        rc = lib.msda_linear256_wgrad(dy.data_ptr(), x.data_ptr(), None, 0, dw.data_ptr(), rows, n_in, n_out,
                                      stream())
        torch.cuda.synchronize()
        ref = dy.double().t() @ x.double()
        print('  rc', rc, lib.msda_last_error() if rc else '', 'rows', rows,
              'rel err ours %.3e torch fp32 %.3e' % (rel(dw, ref), rel(dy.t() @ x, ref)),
              'nan', int(torch.isnan(dw).sum()))
    t_ours = time_ms(lambda: lib.msda_linear256_wgrad(dy.data_ptr(), x.data_ptr(), None, 0, dw.data_ptr(), rows,
                                                      n_in, n_out, stream()))
    t_ref = time_ms(lambda: dy.t() @ x)
    print('  rows %d: tcgen05 3xTF32 %.4f ms   torch fp32 (cuBLAS) %.4f ms' % (rows, t_ours, t_ref))

for width in (128, 256, 1024):
    print('== colsum width=%d' % width)
    rows = 66669
    dy = torch.randn(rows, width, device=dev)
    out = torch.full((width,), float('nan'), device=dev)
    rc = lib.msda_colsum256(dy.data_ptr(), None, out.data_ptr(), rows, width, stream())
    torch.cuda.synchronize()
    ref = dy.double().sum(0)
    print('  rc', rc, 'rel err %.3e  torch %.3e' % (rel(out, ref), rel(dy.sum(0), ref)))
    t_ours = time_ms(lambda: lib.msda_colsum256(dy.data_ptr(), None, out.data_ptr(), rows, width, stream()))
    t_ref = time_ms(lambda: dy.sum(0))
    print('  rows %d: ours %.4f ms   torch %.4f ms' % (rows, t_ours, t_ref))
