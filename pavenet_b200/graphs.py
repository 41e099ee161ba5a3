"""CUDA-graph capture of static-shape stages, forward and backward
("CUDA-graph capture of a whole decoder layer").

A transformer layer around the sampling op is ~100 small kernels forward and
~200 backward; at 300 pose queries or 15 keypoint queries per person each of them
runs for a few microseconds, so the layer's time is the host's launch rate, not the
GPU's.  `GraphedStage` wraps a function of tensors (plus the modules / parameters
it uses) and runs it through `torch.cuda.make_graphed_callables`: one graph launch
for the forward, one for the backward, per distinct input signature.

What makes the attention modules of this package capturable:
  * the C-ABI launchers make no synchronising call (shapes and level starts are
    read on the device, SM count and function attributes are host-side queries);
  * every scratch buffer comes from torch's caching allocator (graph-private pool);
  * dropout inside the GEMM epilogues takes its seed from device memory
    (`functional.device_dropout_seed`), so replays draw new masks; torch's own
    dropout is graph-safe already (Philox offset registered with the graph).

Call `refresh_seed()` once per step before the first graphed stage.
"""
import collections
import contextlib
import itertools

import torch
import torch.nn as nn

from .functional import device_dropout_seed

__all__ = ['GraphedStage', 'refresh_seed', 'quiet_accumulate_grad_stream_warning']

_SEED = {}
_SALT = itertools.count(1)      # one dropout-seed salt per captured graph, process wide


@contextlib.contextmanager
def quiet_accumulate_grad_stream_warning():
    """Capture runs on a side stream, so the parameters' AccumulateGrad nodes (created earlier on
    the default stream) see gradients produced on another stream; autograd inserts the event
    wait it needs and warns -- the wait is intended here.  Silenced for the capture and for the
    backward passes that replay captured stages (clip_model.train_step), not process wide."""
    setter = getattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch', None)
    if setter is None:
        yield
        return
    setter(False)
    try:
        yield
    finally:
        setter(True)


def _seed_tensor(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    t = _SEED.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device='cuda:%d' % key)
        t.random_()
        _SEED[key] = t
    return t


def refresh_seed(device=None):
    """New dropout masks for the next replay of every graphed stage on `device` (one tiny
    kernel, outside the graphs)."""
    _seed_tensor(device if device is not None else torch.cuda.current_device()).random_()


class _Stage(nn.Module):
    """fn(*tensors) with the modules / parameters it touches registered, so the capture
    sees their parameters as graph inputs and returns their gradients."""

    def __init__(self, fn, modules, params, seed, static, salt):
        super().__init__()
        self.mods = nn.ModuleList(modules)
        self.extra = nn.ParameterList(params)
        self._fn, self._seed, self._static, self._salt = fn, seed, static, salt

    def forward(self, *args):
        # the salt keeps the dropout sites of this graph apart from those of every other graph
        # that reads the same per-device seed tensor
        with device_dropout_seed(self._seed, salt=self._salt):
            if self._static is not None:
                return self._fn(*args, static=self._static)
            return self._fn(*args)


class GraphedStage(object):
    """stage = GraphedStage(fn, modules, params); out = stage(*tensors, static=None)

    `static`: an optional hashable passed through to `fn(..., static=static)` for the parts of
    the computation that are Python values (e.g. how many persons each clip holds); it is part
    of the signature, so each distinct value gets its own graph.
    `fn` takes and returns tensors (or tuples of tensors) of fixed shapes for a given
    signature of its inputs, calls nothing that synchronises with the host, and uses only
    the listed modules / parameters.  The first call with a new signature (shapes, dtypes,
    requires_grad, training mode) runs warm-up iterations and captures; later calls replay.
    Not an nn.Module on purpose: it must not re-register the wrapped modules in the model's
    own module tree.

    Every captured graph pins a private memory pool with its saved activations, and a new
    signature costs `warmup_iters` eager runs plus a capture, so the cache is bounded: at most
    `max_graphs` graphs are kept (least recently used evicted), and a signature is only captured
    once it has been seen `capture_after` times -- before that (data-dependent `static` keys that
    rarely repeat, e.g. per-clip person counts) the stage runs eagerly."""

    def __init__(self, fn, modules=(), params=(), warmup_iters=3, max_graphs=8, capture_after=1):
        self.fn, self.modules, self.params = fn, list(modules), list(params)
        self.warmup_iters = warmup_iters
        self.max_graphs, self.capture_after = max(1, int(max_graphs)), max(1, int(capture_after))
        self._graphs = collections.OrderedDict()
        self._seen = collections.Counter()
        self.stats = {'captures': 0, 'evictions': 0, 'eager_calls': 0, 'replays': 0,
                      'library_launches_replayed': 0}
        self._launches = {}           # signature -> library kernel launches per forward+backward replay

    def _signature(self, args, static):
        training = tuple(m.training for m in self.modules)
        return (static,) + training + tuple((tuple(a.shape), a.dtype, a.requires_grad, a.device.index) for a in args)

    def captured_signatures(self):
        return list(self._graphs)

    def __call__(self, *args, static=None):
        for a in args:
            if not (isinstance(a, torch.Tensor) and a.is_cuda):
                raise RuntimeError('GraphedStage takes CUDA tensors only, got %r' % (type(a),))
        key = self._signature(args, static)
        graphed = self._graphs.get(key)
        if graphed is not None:
            self._graphs.move_to_end(key)
            self.stats['replays'] += 1
            self.stats['library_launches_replayed'] += self._launches.get(key, 0)
            return graphed(*args)
        self._seen[key] += 1
        if len(self._seen) > 64 * self.max_graphs:      # the counter itself stays bounded
            self._seen = collections.Counter({k: v for k, v in self._seen.most_common(8 * self.max_graphs)})
        if self._seen[key] < self.capture_after:
            self.stats['eager_calls'] += 1
            if static is not None:
                return self.fn(*args, static=static)
            return self.fn(*args)
        while len(self._graphs) >= self.max_graphs:
            old, _ = self._graphs.popitem(last=False)    # frees that graph's memory pool
            self._launches.pop(old, None)
            self.stats['evictions'] += 1
        stage = _Stage(self.fn, self.modules, self.params, _seed_tensor(args[0].device), static,
                       next(_SALT))
        # make_graphed_callables keys its replay on the wrapper's training flag; the wrapped
        # modules keep their own modes (frozen BatchNorm stays in eval)
        sample = tuple(a.detach().clone().requires_grad_(a.requires_grad) for a in args)
        from . import _capi
        before = _capi.launch_count()
        with quiet_accumulate_grad_stream_warning():
            graphed = torch.cuda.make_graphed_callables(stage, sample, num_warmup_iters=self.warmup_iters,
                                                        allow_unused_input=True)
        # the warm-up iterations and the capture each ran the stage's forward + backward once
        self._launches[key] = (_capi.launch_count() - before) // (self.warmup_iters + 1)
        self._graphs[key] = graphed
        self.stats['captures'] += 1
        return graphed(*args)
