"""Host<->device link bandwidth of this box: each direction alone, and both at once — for one
GPU, or for all ranks of a torchrun launch AT THE SAME TIME (what the N-GPU `e2e` leg of bench.py
does to the host: N x 239 MB up and N x 239 MB down per step through one host memory system).

    python tools/pcie_duplex.py
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_duplex.py [--bind]

--bind: pin each rank to the CPUs NVML reports as local to its GPU before allocating the pinned
buffers (first touch then places them on that NUMA node), as bench.py does for its e2e leg.
Pinned host memory, cudaMemcpyAsync on two streams; wall clock between barriers.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MB = 1 << 20


def run(nbytes, up, down, barrier, reps=8):
    h_up = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_up = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    d_dn = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        barrier()
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
    import bench
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    affinity = bench.bind_to_gpu_numa(local) if '--bind' in sys.argv else {'bound': False}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        barrier = dist.barrier
    else:
        barrier = lambda: None     # noqa: E731
    rows = []
    for mb in (12, 239):
        n = mb * MB
        rows.append((mb, run(n, True, False, barrier), run(n, False, True, barrier),
                     run(n, True, True, barrier)))
    for r in range(world):
        barrier()
        if r == rank:
            for mb, u, d, x in rows:
                print('rank %d/%d  %4d MiB  H2D alone %5.1f GB/s   D2H alone %5.1f GB/s   duplex %5.1f GB/s each way   %s'
                      % (rank, world, mb, u, d, x, affinity), flush=True)
    if world > 1:
        dist.destroy_process_group()
