"""Effective host<->device link rate of QUEUED back-to-back cudaMemcpyAsync calls as a function of the
copy size: H2D alone, D2H alone, and both directions at once on two streams (what the host-buffer
pipeline of msda_forward_backward_host issues).  CUDA events around the whole queue, 240 MB per direction."""
import sys

import torch

MB = 1 << 20


def run(size, up, down, total=240 * MB):
    n = max(1, total // size)
    h_up = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(n * size, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(n * size, dtype=torch.uint8, device='cuda')
    d_dn = torch.empty(n * size, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        if up:
            with torch.cuda.stream(s1):
                for i in range(n):
                    d_up[i * size:(i + 1) * size].copy_(h_up[i * size:(i + 1) * size], non_blocking=True)
                e1.record()
        if down:
            with torch.cuda.stream(s2):
                for i in range(n):
                    h_dn[i * size:(i + 1) * size].copy_(d_dn[i * size:(i + 1) * size], non_blocking=True)
                e2.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1) if up else 0.0, e0.elapsed_time(e2) if down else 0.0)
        best = min(best, ms)
    return n * size / (best * 1e-3) / 1e9


if __name__ == '__main__':
    print('copy size    H2D alone   D2H alone   duplex (each way)   [GB/s, queued back-to-back copies]')
    for mb in (1, 2.25, 4.5, 11.4, 22.8, 60, 240):
        size = int(mb * MB) // 256 * 256
        print('%6.2f MiB   %7.1f     %7.1f     %7.1f' % (mb, run(size, True, False), run(size, False, True), run(size, True, True)), flush=True)
