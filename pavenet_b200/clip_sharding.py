"""Clip-sharded data parallelism for the op: the only multi-GPU structure the
path has.

Every kernel thread touches only its own batch entry (one frame, or one clip
in the fused multi-frame view), so clips are independent units: each rank
runs its own clips, with NO collective inside the op, forward or backward.
The reference trains the same way — one clip per GPU under DDP
(configs/_base_/datasets/posetrack17_video_keypoint.py:88,
opera/apis/train.py:153-162); its only collectives are the parameter-gradient
all-reduce and a few scalar reductions, all outside this op.

This module holds the host-side bookkeeping: which clips a rank owns, and the
max-over-ranks reduction used to time a sharded run.
"""
import torch
import torch.distributed as dist

__all__ = ['shard_range', 'shard_sizes', 'max_over_ranks', 'sum_over_ranks']


def shard_sizes(num_clips, world_size):
    """Clips per rank: as even as possible, earlier ranks take the remainder."""
    if world_size <= 0:
        raise ValueError('world_size must be positive, got %r' % (world_size,))
    if num_clips < 0:
        raise ValueError('num_clips must be non-negative, got %r' % (num_clips,))
    base, extra = divmod(num_clips, world_size)
    return [base + (1 if r < extra else 0) for r in range(world_size)]


def shard_range(num_clips, rank, world_size):
    """[begin, end) of the clips owned by `rank`; contiguous, disjoint, covering."""
    if not 0 <= rank < world_size:
        raise ValueError('rank %r out of range for world_size %r' % (rank, world_size))
    sizes = shard_sizes(num_clips, world_size)
    begin = sum(sizes[:rank])
    return begin, begin + sizes[rank]


def _reduce(value, op, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device='cpu'):
    """Max of a per-rank scalar (elapsed device time) over all ranks."""
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device='cpu'):
    """Sum of a per-rank scalar (units processed) over all ranks."""
    return _reduce(value, dist.ReduceOp.SUM, device)
