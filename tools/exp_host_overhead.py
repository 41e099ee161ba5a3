"""Where does the host time of one eager op call go?  (PETR shape; the GPU work is ~15 us per kernel.)"""
import sys
import time

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from pavenet_b200 import _capi, functional as F  # noqa: E402

dev = torch.device('cuda', 0)
p = bench.make_problem('petr_cfg1', seed=0, device=dev)
lib = _capi.load()
gv = torch.empty_like(p['value']); gl = torch.empty_like(p['loc']); ga = torch.empty_like(p['aw'])
N = 2000


def timeit(name, fn, sync_every=200):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        fn()
        if (i + 1) % sync_every == 0:
            torch.cuda.synchronize()      # keep the launch queue from filling up (that would measure the GPU)
    dt = (time.perf_counter() - t0) / N * 1e6
    torch.cuda.synchronize()
    print('%-58s %7.2f us' % (name, dt), flush=True)


args = (p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'])
timeit('ms_deform_attn_forward (whole wrapper)', lambda: F.ms_deform_attn_forward(*args, 64), 50)
timeit('ms_deform_attn_forward with clear=', lambda: F.ms_deform_attn_forward(*args, 64, clear=gv), 50)
timeit('ms_deform_attn_backward (whole wrapper)', lambda: F.ms_deform_attn_backward(*args, p['grad_out'], gv, gl, ga, 64), 50)
timeit('_check_inputs', lambda: F._check_inputs(*args, 64))
timeit('torch.empty(out)', lambda: torch.empty((1, 300, 256), dtype=torch.float32, device=dev))
def ctx():
    with torch.cuda.device(dev):
        pass
timeit('with torch.cuda.device(dev)', ctx)
timeit('torch.cuda.current_stream().cuda_stream', lambda: torch.cuda.current_stream().cuda_stream)
timeit('6 x data_ptr()', lambda: [t.data_ptr() for t in (args + (gv,))])
d = p['dims']
out = torch.empty((1, 300, 256), device=dev)
ptrs = [t.data_ptr() for t in args] + [out.data_ptr()]
stream = torch.cuda.current_stream().cuda_stream
timeit('lib.msda_forward_clear (ctypes + C + launch)', lambda: lib.msda_forward_clear(
    *ptrs, d['B'], d['S'], d['M'], d['D'], d['L'], d['Q'], d['P'], 0, 0, None, 0, stream), 50)
timeit('lib.msda_launch_count (empty ctypes call)', lambda: lib.msda_launch_count())
timeit('gv.zero_()', lambda: gv.zero_(), 50)
