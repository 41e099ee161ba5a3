"""Are the small-Q workloads bound by the host (Python + ctypes launch cost) rather than by the GPU?
For each workload: (a) eager loop as bench.py runs it, device time per step and host time per step (the
host loop timed without synchronising); (b) the same steps captured into one CUDA graph (4 input
sets) and replayed.  Usage: python tools/exp_launch_bound.py [workload ...]"""
import sys
import time

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from pavenet_b200.functional import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402


def main():
    wls = sys.argv[1:] or ['petr_cfg1', 'pose_cfg3_t3', 'pose_cfg3', 'encoder_cfg2']
    dev = torch.device('cuda', 0)
    for wl in wls:
        sets = 4
        probs = [bench.make_problem(wl, seed=i, device=dev) for i in range(sets)]
        bufs = [dict(gv=torch.empty_like(p['value']), gl=torch.empty_like(p['loc']), ga=torch.empty_like(p['aw']))
                for p in probs]
        fold = bench.WORKLOADS[wl]['kind'] == 'pose'

        def step(i):
            p, b = probs[i % sets], bufs[i % sets]
            out = ms_deform_attn_forward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], 64,
                                         clear=b['gv'] if fold else None)
            if not fold:
                b['gv'].zero_()
            ms_deform_attn_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], p['grad_out'],
                                    b['gv'], b['gl'], b['ga'], 64)
            return out

        for i in range(20):
            step(i)
        torch.cuda.synchronize()
        n = 400
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n):
            step(i)
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        eager_ms = e0.elapsed_time(e1) / n

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(8):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(sets):
                step(i)
        for _ in range(5):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n // sets):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / n
        print('%-14s eager %.4f ms/step (host loop %.4f ms/step)   graph replay %.4f ms/step   ratio %.2f'
              % (wl, eager_ms, t_host * 1e3 / n, graph_ms, eager_ms / graph_ms), flush=True)


if __name__ == '__main__':
    main()
