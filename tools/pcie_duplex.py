"""Host<->device link bandwidth of this box: each direction alone, and both at once.

The end-to-end (`e2e`) bench number moves 239 MB up and 239 MB down per step; this tells how
close its 6.5 ms is to what the link can do.  Pinned host memory, cudaMemcpyAsync on two streams.
"""
import time
import torch

MB = 1 << 20


def run(nbytes, up, down, reps=10):
    h_up = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    d_dn = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9


if __name__ == '__main__':
    for mb in (12, 64, 239):
        n = mb * MB
        print('%4d MiB  H2D alone %5.1f GB/s   D2H alone %5.1f GB/s   duplex %5.1f GB/s each way' %
              (mb, run(n, True, False), run(n, False, True), run(n, True, True)), flush=True)
