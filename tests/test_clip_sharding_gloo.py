"""N > 1 host-side logic on CPU: world_size-2 gloo.  The op has no data-path
collective (clips are independent), so what is tested is the bookkeeping the
sharded bench relies on: disjoint covering clip ranges, and the MAX / SUM
reductions used for device-time and throughput."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pavenet_b200 import clip_sharding


def test_shard_ranges_cover_and_are_disjoint():
    for clips in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            sizes = clip_sharding.shard_sizes(clips, world)
            assert sum(sizes) == clips and max(sizes) - min(sizes) <= 1
            covered = []
            for r in range(world):
                b, e = clip_sharding.shard_range(clips, r, world)
                covered.extend(range(b, e))
            assert covered == list(range(clips))
    with pytest.raises(ValueError):
        clip_sharding.shard_range(4, 2, 2)
    with pytest.raises(ValueError):
        clip_sharding.shard_sizes(4, 0)


def test_reductions_without_process_group():
    assert clip_sharding.max_over_ranks(3.5) == 3.5
    assert clip_sharding.sum_over_ranks(2) == 2.0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, clips, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        b, e = clip_sharding.shard_range(clips, rank, world)
        # each rank "processes" its clips: a per-clip value only it computes
        mine = torch.zeros(clips, dtype=torch.float64)
        for c in range(b, e):
            mine[c] = (c + 1) ** 2
        elapsed = 10.0 + rank                     # rank 1 is the slow one
        t_max = clip_sharding.max_over_ranks(elapsed)
        n_sum = clip_sharding.sum_over_ranks(e - b)
        # the only "exchange" of the path is at the very end (metrics); data
        # never moves between ranks — verify ownership by gathering for the test
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        torch.save(dict(t_max=t_max, n_sum=n_sum, gathered=torch.stack(gathered), rng=(b, e)),
                   os.path.join(out_dir, 'r%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world, clips = 2, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, clips, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r)) for r in range(world)]
    for r in res:
        assert r['t_max'] == 11.0            # max over ranks, not this rank's time
        assert r['n_sum'] == clips
    assert res[0]['rng'] == (0, 3) and res[1]['rng'] == (3, 5)
    g = res[0]['gathered']
    # every clip was produced by exactly one rank
    assert ((g != 0).sum(0) == 1).all()
    assert torch.equal(g.sum(0), torch.arange(1, clips + 1, dtype=torch.float64) ** 2)


def _flat_worker(rank, world, port, out_dir):
    """The gradient exchange of the clip-sharded training step: every .grad is a view into one
    flat bucket, one all-reduce averages it (clip_model.FlatGradients)."""
    from pavenet_b200 import clip_model
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                               # same weights on every rank
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 2))
        net[0].bias.requires_grad_(False)                  # frozen parameters stay out of the bucket
        flat = clip_model.FlatGradients(net)
        x = torch.full((3, 6), float(rank + 1))            # each rank sees its own clip
        for _ in range(2):                                 # second pass: zero() really clears the views
            flat.zero()
            net(x).square().sum().backward()
            local = flat.flat.clone()
            flat.all_reduce_mean()
        views_ok = all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in flat.params)
        torch.save(dict(local=local, reduced=flat.flat.clone(), views_ok=views_ok,
                        n=sum(p.numel() for p in flat.params)), os.path.join(out_dir, 'f%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_bucket_all_reduce_gloo(tmp_path):
    world = 2
    mp.spawn(_flat_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), 'f%d.pt' % r)) for r in range(world)]
    mean = (res[0]['local'] + res[1]['local']) / 2
    for r in res:
        assert r['views_ok'] and r['n'] == 6 * 5 + 5 * 2 + 2 == r['reduced'].numel()
        assert torch.allclose(r['reduced'], mean, rtol=1e-6, atol=1e-7)
    assert not torch.allclose(res[0]['local'], res[1]['local'])


class _ThreeStageNet(torch.nn.Module):
    """backbone -> encoder -> head, with the stage-boundary hooks of clip_model.PaveNetR50."""
    grad_exchange = None

    def __init__(self):
        super().__init__()
        self.backbone = torch.nn.Linear(6, 8)
        self.encoder = torch.nn.Linear(8, 8)
        self.head = torch.nn.Linear(8, 3)
        self.launch_log = []

    def gradient_buckets(self):
        return [list(self.head.parameters()), list(self.encoder.parameters()),
                list(self.backbone.parameters())]

    def _hook(self, tensor, n):
        ex = self.grad_exchange

        def hook(grad):
            self.launch_log.append((n, ex._launched))
            ex.launch_through(n)
        tensor.register_hook(hook)

    def forward(self, x):
        f = torch.relu(self.backbone(x))
        self._hook(f, 2)                  # backward reaches the backbone: head + encoder done
        e = torch.relu(self.encoder(f))
        self._hook(e, 1)                  # backward reaches the encoder: head done
        return self.head(e)


def _bucket_worker(rank, world, port, out_dir):
    """Stage buckets all-reduced from autograd hooks while the backward is still running
    (clip_model.FlatGradients(overlap=True)); the result must equal the plain mean."""
    from pavenet_b200 import clip_model
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = _ThreeStageNet()
        flat = clip_model.FlatGradients(net, overlap=True)
        assert net.grad_exchange is flat and len(flat.ranges) == 3
        x = torch.randn(4, 6) * (rank + 1)
        for _ in range(2):
            flat.zero()
            net.launch_log.clear()
            net(x).square().sum().backward()
            launched_in_backward = flat._launched
            flat.all_reduce_mean()
        ref = _ThreeStageNet()
        ref.load_state_dict(net.state_dict())
        ref.grad_exchange = type('No', (), {'_launched': 0, 'launch_through': lambda self, n: None})()
        ref(x).square().sum().backward()
        local = torch.cat([p.grad.reshape(-1) for b in ref.gradient_buckets() for p in b])
        torch.save(dict(local=local, reduced=flat.flat.clone(), log=list(net.launch_log),
                        launched_in_backward=launched_in_backward), os.path.join(out_dir, 'b%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_stage_buckets_overlap_gradient_exchange_gloo(tmp_path):
    world = 2
    mp.spawn(_bucket_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), 'b%d.pt' % r)) for r in range(world)]
    mean = (res[0]['local'] + res[1]['local']) / 2
    for r in res:
        assert torch.allclose(r['reduced'], mean, rtol=1e-6, atol=1e-7)
        # the head's bucket went out when the backward reached the encoder, the encoder's when it
        # reached the backbone; only the backbone's bucket was left for after the backward
        assert r['log'] == [(1, 0), (2, 1)] and r['launched_in_backward'] == 2
