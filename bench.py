#!/usr/bin/env python
"""bench.py — throughput of the multi-scale deformable attention hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload encoder_cfg2|pose_cfg3|pose_cfg3_t3|petr_cfg1|stress_cfg5|stress_cfg5_big]
                    [--value-dtype f32|bf16]

A "step" is one forward + backward pass of the op over one clip of synthetic,
PAVE-Net-shaped input (BASELINE.json configs[1] by default: spatial-encoder
attention, R-50 features of a 3-frame clip at 800x1333, 8 heads x 4 levels x 4
points, 66 669 queries).  Metric: queries/s (whole job, all GPUs).

Prints ONE JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (backward): algorithmic bytes / CUDA-event
                duration against the measured HBM copy peak
  roofline_fwd / roofline_step   the same for the forward kernel and for the
                whole step (fwd + grad_value zero-fill + bwd)
  roofline_onchip   the limiter that actually binds: 128-byte value rows gathered through
                L1 (forward) / reduced into L2 (backward) per second against the
                micro-benchmarked peak of that instruction stream
  cpu_baseline  the reference's CPU path (per-level grid_sample + autograd),
                restated in oracle/msda_oracle.py, timed on this host's cores
  e2e           the same metric through the C-ABI host-buffer call
                (msda_forward_backward_host) with pinned HOST buffers: H2D of
                the inputs and D2H of output and all gradients inside the timed
                region, pipelined inside the library
  e2e_autograd  the same copies issued serially around the autograd Function
  launch        how the timed steps were launched.  The small-Q (pose) workloads run 40-150 us of
                kernels per step, less than the Python + ctypes cost of launching them one by one, so
                their K steps are ALSO timed as replays of a CUDA graph holding `sets` steps (what the
                library's own GraphedStage does for the pose decoder); `value` / `ms_per_step` then come
                from that leg, `ms_per_step_eager` and the per-kernel event times from the eager leg
Multi-GPU: clips are independent, so each rank runs its own clips (weak
scaling) with no collective on the data path; one all-reduce(MAX) of the
elapsed time at the end.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R50_LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]       # 800x1333, strides 8/16/32/64
BIG_LEVELS = [(150, 250), (75, 125), (38, 63), (19, 32)]      # 1200x2000

WORKLOADS = {
    # name: (description, frames-as-batch B, fused frames T, queries, points, levels, kind)
    'encoder_cfg2': dict(desc='PAVE-Net spatial-encoder MSDA fwd+bwd, R-50 800x1333, 3-frame clip, '
                              '8 heads x 4 levels x 4 points', B=3, T=1, Q=None, P=4,
                         levels=R50_LEVELS, kind='encoder'),
    'pose_cfg3': dict(desc='pose-decoder pose-aware attention fused over T=5 frames, 300 pose queries '
                           'x 17 keypoints, R-50 800x1333', B=1, T=5, Q=300, P=17,
                      levels=R50_LEVELS, kind='pose'),
    'pose_cfg3_t3': dict(desc='pose-decoder pose-aware attention fused over T=3 frames, 300 pose '
                              'queries x 15 keypoints (PoseTrack config)', B=1, T=3, Q=300, P=15,
                         levels=R50_LEVELS, kind='pose'),
    'petr_cfg1': dict(desc='PETR pose attention, 1 frame, 300 queries x 17 keypoints', B=1, T=1,
                      Q=300, P=17, levels=R50_LEVELS, kind='pose'),
    'encoder_cfg2_rand': dict(desc='config 2 with ADVERSARIAL locations: every sample uniform in the image (no spatial '
                                   'coherence between neighbouring queries; SURVEY.md section 8d)', B=3, T=1, Q=None,
                              P=4, levels=R50_LEVELS, kind='encoder', rand_loc=True),
    'stress_cfg5': dict(desc='encoder stress: 8 frames at 800x1333, 4 levels x 4 points', B=8, T=1,
                        Q=None, P=4, levels=R50_LEVELS, kind='encoder'),
    'stress_cfg5_big': dict(desc='encoder stress: 8 frames at 1200x2000 (Swin-L high-res), 4 levels x 4 points',
                            B=8, T=1, Q=None, P=4, levels=BIG_LEVELS, kind='encoder'),
}
M_HEADS, D_HEAD = 8, 32
MODEL_WORKLOADS = ('pavenet_step',)


# ---------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------
def make_problem(wl, seed, device, value_dtype=torch.float32, frames=None):
    """Seeded synthetic inputs of one step, generated on `device`.

    encoder: queries are the pixels of all levels; locations = reference grid
    + the module's ring-offset init (multi_scale_deform_attn.py:286-297) + N(0,1)
    pixels of noise — spatially coherent, like a trained encoder.
    pose: per-query pose box (centre ~U(0.1,0.9), size ~U(0.05,0.4)), K keypoints
    uniform in the box + N(0,0.02) offsets, small drift between frames.
    """
    cfg = WORKLOADS[wl]
    g = torch.Generator(device=device).manual_seed(seed)
    levels = cfg['levels']
    T, P = cfg['T'], cfg['P']
    B = frames if frames is not None else cfg['B']
    L = len(levels)
    S = sum(h * w for h, w in levels)
    M, D = M_HEADS, D_HEAD
    shapes = torch.tensor(levels, dtype=torch.int64, device=device)
    sizes = shapes[:, 0] * shapes[:, 1]
    lsi = torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])

    def randn(*s):
        return torch.randn(*s, generator=g, device=device)

    def rand(*s):
        return torch.rand(*s, generator=g, device=device)

    if cfg['kind'] == 'encoder':
        Q = S
        ref = []
        for h, w in levels:
            ys = (torch.arange(h, device=device, dtype=torch.float32) + 0.5) / h
            xs = (torch.arange(w, device=device, dtype=torch.float32) + 0.5) / w
            yy, xx = torch.meshgrid(ys, xs, indexing='ij')
            ref.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
        ref = torch.cat(ref)                                            # (S, 2)
        th = torch.arange(M, device=device, dtype=torch.float32) * (2.0 * torch.pi / M)
        ring = torch.stack([th.cos(), th.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        steps = torch.arange(1, P + 1, device=device, dtype=torch.float32)
        off = ring[:, None, None, :] * steps[None, None, :, None]          # (M,1,P,2)
        norm = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32, device=device)
        loc = ref[None, :, None, None, None, :] + (
            off[None, None] + randn(B, Q, M, L, P, 2)) / norm[None, None, None, :, None, :]
        if cfg.get('rand_loc'):
            loc = rand(B, Q, M, L, P, 2)
        Lk = L
        # experiment: hand the queries to the op in patch order (PW x PH pixel patches inside each
        # level) instead of raster order, to measure what a patch-shaped block would gain in L1 hits
        order = os.environ.get('PAVENET_BENCH_QUERY_ORDER', '')
        if order.startswith('patch'):
            pw, ph = (int(v) for v in order[5:].split('x'))
            perm, base = [], 0
            for h, w in levels:
                yy, xx = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device),
                                        indexing='ij')
                key = (((yy // ph) * ((w + pw - 1) // pw) + xx // pw) * ph + yy % ph) * pw + xx % pw
                perm.append(base + key.reshape(-1).argsort())
                base += h * w
            loc = loc[:, torch.cat(perm)]
    else:
        Q = cfg['Q']
        Lk = T * L
        centre = rand(B, Q, 1, 1, 1, 2) * 0.8 + 0.1
        size = rand(B, Q, 1, 1, 1, 2) * 0.35 + 0.05
        kpt = centre + (rand(B, Q, 1, 1, P, 2) - 0.5) * size                 # shared by levels
        drift = randn(B, Q, 1, T, 1, 1, 2) * 0.01                             # per frame
        loc = (kpt[:, :, :, None] + drift).expand(B, Q, 1, T, L, P, 2)
        loc = loc.reshape(B, Q, 1, Lk, P, 2) + randn(B, Q, M, Lk, P, 2) * 0.02
        shapes = shapes.repeat(T, 1)
        lsi = (torch.arange(T, device=device)[:, None] * S + lsi[None, :]).reshape(-1)
    value = randn(B, T * S, M, D).to(value_dtype)
    aw = torch.softmax(randn(B, Q, M, Lk * P), -1).view(B, Q, M, Lk, P)
    grad_out = randn(B, Q, M * D)
    return dict(value=value.contiguous(), shapes=shapes.contiguous(), lsi=lsi.contiguous(),
                loc=loc.contiguous(), aw=aw.contiguous(), grad_out=grad_out.contiguous(),
                dims=dict(B=B, S=T * S, M=M, D=D, L=Lk, Q=Q, P=P))


def algorithmic_bytes(dims, value_bytes=4, grad_value_bytes=4):
    """unique bytes a perfect kernel must move.
    fwd = V + LOC + W + O;  bwd = V + LOC + W + O(grad_out) + Vg + LOC(grad) + W(grad)
    with V = min(value bytes, bytes actually sampled)."""
    B, S, M, D, L, Q, P = (dims[k] for k in 'BSMDLQP')
    n_s = B * Q * M * L * P
    V = min(B * S * M * D * value_bytes, n_s * 4 * D * value_bytes)
    Vg = min(B * S * M * D * grad_value_bytes, n_s * 4 * D * grad_value_bytes)
    LOC, W, O = 8 * n_s, 4 * n_s, 4 * B * Q * M * D
    return dict(fwd=V + LOC + W + O, bwd=V + LOC + W + O + Vg + LOC + W)


# ---------------------------------------------------------------------------
# clocks (NVML) sampled during the timed region
# ---------------------------------------------------------------------------
class ClockSampler(object):
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # noqa: BLE001 - clocks are diagnostics, never fatal
            self._nv = None
            self.error = str(exc)

    _REASONS = (('hw_slowdown', 0x8), ('sw_power_cap', 0x4), ('sw_thermal_slowdown', 0x20),
                ('hw_thermal_slowdown', 0x40), ('hw_power_brake_slowdown', 0x80),
                ('sync_boost', 0x10), ('applications_clocks_setting', 0x2))

    def _once(self):
        nv = self._nv
        self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:  # noqa: BLE001
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for name, bit in self._REASONS:
            if bits & bit:
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._once()
            except Exception:  # noqa: BLE001
                return
            self._stop.wait(0.005)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [],
                    'note': getattr(self, 'error', 'no samples')}
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ---------------------------------------------------------------------------
# CPU baseline: the reference's grid_sample path, restated in oracle/
# ---------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every
    core this process may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def bind_to_gpu_numa(index):
    """Restrict this process to the CPUs NVML reports as local to GPU `index`, so that pinned
    host buffers allocated afterwards are first-touched on that GPU's NUMA node.  Returns what
    was done (for the JSON line); never fatal."""
    info = {'bound': False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info.update(gpu_local_cpus=len(cpus), allowed_cpus=len(allowed), usable=len(use))
        try:
            info['numa_node'] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:  # noqa: BLE001 - older NVML
            pass
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info['bound'] = True
    except Exception as exc:  # noqa: BLE001
        info['note'] = '%s: %s' % (type(exc).__name__, exc)
    return info


def cpu_reference_step(prob_cpu):
    """One fwd+bwd of the reference CPU algorithm on CPU tensors; returns seconds."""
    from oracle import msda_oracle as O
    v = prob_cpu['value'].float().requires_grad_()
    loc = prob_cpu['loc'].clone().requires_grad_()
    aw = prob_cpu['aw'].clone().requires_grad_()
    t0 = time.perf_counter()
    out = O.grid_sample_port(v, prob_cpu['shapes'], loc, aw)
    out.backward(prob_cpu['grad_out'])
    return time.perf_counter() - t0


def cpu_sample_problem(wl, frames=1, queries=None, seed=1234):
    """A bounded sample of the workload for the CPU legs: `frames` batch
    entries and (optionally) the first `queries` queries of each."""
    prob = make_problem(wl, seed=seed, device='cpu', frames=frames)
    if queries is not None and queries < prob['dims']['Q']:
        for k in ('loc', 'aw', 'grad_out'):
            prob[k] = prob[k][:, :queries].contiguous()
        prob['dims']['Q'] = queries
    return prob


def measure_cpu_baseline(wl, budget_s=20.0):
    use_all_host_threads()
    cfg = WORKLOADS[wl]
    full_q = sum(h * w for h, w in cfg['levels']) if cfg['kind'] == 'encoder' else cfg['Q']
    prob = cpu_sample_problem(wl, frames=1, queries=min(full_q, 4096))
    t = cpu_reference_step(prob)                       # warm-up + calibration
    per_query = t / prob['dims']['Q']
    q = int(max(64, min(full_q, budget_s / 4 / max(per_query, 1e-9))))
    prob = cpu_sample_problem(wl, frames=1, queries=q)
    cpu_reference_step(prob)
    best = min(cpu_reference_step(prob) for _ in range(3))
    return {'value': prob['dims']['Q'] / best, 'unit': 'queries/s',
            'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '1 batch entry x %d of %d queries of %s, fwd+bwd, fp32, best of 3 after '
                      '1 warm-up (oracle.grid_sample_port + autograd)' % (q, full_q, wl),
            'ms_per_sample': best * 1e3}


# ---------------------------------------------------------------------------
_JSON_OUT = None


def protect_stdout():
    """The caller reads ONE JSON line from stdout; native libraries write there too (NCCL prints
    its version banner to fd 1 at NCCL_DEBUG >= VERSION).  Keep a private duplicate of stdout for
    the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(text):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(text + '\n')
    out.flush()


def dist_setup(gpus):
    # NCCL announces its version on STDOUT at NCCL_DEBUG=VERSION, which would precede the
    # one JSON line this script owes its caller
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
        os.environ['NCCL_DEBUG'] = 'WARN'
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return world, rank, local


def workload_config(wl, dims, sets, world):
    """The `config` object of an op workload -- built by ONE function for both arms (ours and
    --impl reference), so that the two lines describe the same job key for key."""
    cfg = WORKLOADS[wl]
    return {'workload': wl, 'description': cfg['desc'], 'dims': dict(dims),
            'queries_per_step': dims['B'] * dims['Q'],
            'l2_policy': 'inputs larger than L2: %d distinct clips rotated per step' % sets,
            'parallelism': 'clip-sharded x%d, no data-path collective' % world}


def run_reference_arm(args, world, rank):
    """--impl reference: the reference's CPU implementation of the path (the
    oracle port; the Python reference itself cannot travel to the GPU box),
    all host threads, rank 0 only.

    Same job as our arm: every step is one fwd+bwd over ALL batch entries and ALL queries of
    the workload, `--sets` distinct clips rotated.  Only if that would take more than ~5 minutes
    for steps + warmup on this host is a step cut down to a bounded sample (first batch
    entries / first queries), and the line says so (`cpu_baseline.sample`)."""
    if rank != 0:
        return
    cores = use_all_host_threads()
    wl = args.workload
    cfg = WORKLOADS[wl]
    full_q = sum(h * w for h, w in cfg['levels']) if cfg['kind'] == 'encoder' else cfg['Q']
    full_b = args.frames or cfg['B']
    probe = cpu_sample_problem(wl, frames=1, queries=min(full_q, 2048))
    cpu_reference_step(probe)
    t = min(cpu_reference_step(probe) for _ in range(2))
    per_query = t / probe['dims']['Q']
    n_steps = max(1, args.steps + args.warmup)
    est_full = per_query * full_q * full_b * n_steps
    if est_full <= 300.0:
        frames, q = full_b, full_q
    else:                                    # bounded sample: ~2 minutes in all
        per_step = 120.0 / n_steps
        frames = max(1, min(full_b, int(per_step / (per_query * full_q))))
        q = full_q if frames > 1 or per_query * full_q <= per_step else \
            int(max(32, per_step / per_query))
    n_sets = max(1, min(args.sets, 4 if frames * q * 16 * 8 * 12 < 2e8 else 2))
    probs = [cpu_sample_problem(wl, frames=frames, queries=q, seed=1234 + i) for i in range(n_sets)]
    for i in range(args.warmup):
        cpu_reference_step(probs[i % n_sets])
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_reference_step(probs[i % n_sets])
    dt = time.perf_counter() - t0
    q_step = frames * q
    qps = q_step * args.steps / dt
    full = frames == full_b and q == full_q
    sample = ('the full workload: %d batch entries x %d queries per step' % (frames, q) if full else
              'bounded sample: %d of %d batch entries x %d of %d queries per step'
              % (frames, full_b, q, full_q))
    dims = dict(probs[0]['dims'], B=full_b, Q=full_q)
    line = {
        'impl': 'reference', 'metric': 'deform-attn fwd+bwd queries/s', 'value': qps,
        'unit': 'queries/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl, dims, args.sets, 1),
        'same_job_as_ours': full, 'queries_timed_per_step': q_step,
        'cpu_baseline': {'value': qps, 'unit': 'queries/s', 'cores': cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': qps, 'unit': 'queries/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(json.dumps(line))


def measure_model_step(args, world, rank, local, steps, warmup):
    """BASELINE config 4: clip-sharded PAVE-Net R-50 training step (1 clip per GPU as in the
    reference, configs/_base_/datasets/posetrack17_video_keypoint.py:88): forward + backward +
    NCCL gradient all-reduce (bucketed by stage and overlapped with the backward,
    clip_model.FlatGradients) + grad-clip + AdamW.  The process group must exist already when
    world > 1.  Returns the record on every rank (timings are max over ranks)."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from pavenet_b200 import _capi, clip_model, clip_sharding
    device = torch.device('cuda', local)
    torch.manual_seed(0)
    vdt = None if args.value_dtype == 'f32' else torch.bfloat16
    model = clip_model.PaveNetR50(value_dtype=vdt).to(device).train()
    model.fold_frozen_bn = os.environ.get('PAVENET_FOLD_BN', '1') != '0'   # A/B switch, default on
    if args.graphs:
        model.enable_graphs()
    ddp = flat = None
    if args.grad_exchange == 'ddp':
        ddp = DDP(model, device_ids=[local], broadcast_buffers=False) if world > 1 else None
    else:
        if world > 1:                                  # same initial weights on every rank
            for p in model.parameters():
                dist.broadcast(p.data, 0)
        flat = clip_model.FlatGradients(model, overlap=args.grad_exchange == 'overlap')
    opt = clip_model.build_optimizer(model)
    clips_per_gpu = 1
    batches = [clip_model.synthetic_clip_batch(clips_per_gpu, device, seed=100 * rank + i)
               for i in range(2)]
    # two alternating synthetic batches; a graphed stage captures a signature the second time
    # it sees it, so every graph exists after 2 x 2 steps (+ the capture's own warm-up runs)
    for i in range(max(warmup, 6)):
        clip_model.train_step(model, opt, *batches[i % 2], ddp_model=ddp, flat_grads=flat)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = _capi.launch_count() + model.library_launches_replayed()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exposed = []
    e0.record()
    for i in range(steps):
        loss = clip_model.train_step(model, opt, *batches[i % 2], ddp_model=ddp, flat_grads=flat)
        if flat is not None and flat.exposed_events is not None:
            exposed.append(flat.exposed_events)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clock_info = clocks.stop()
    ms = clip_sharding.max_over_ranks(e0.elapsed_time(e1), device)
    clips = clip_sharding.sum_over_ranks(clips_per_gpu * steps, device)
    exposed_ms = (sum(a.elapsed_time(b) for a, b in exposed) / len(exposed)) if exposed else 0.0
    exposed_ms = clip_sharding.max_over_ranks(exposed_ms, device)
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    rec = {
        'metric': 'PAVE-Net R-50 training clips/s', 'value': clips / (ms * 1e-3), 'unit': 'clips/s',
        'n_gpus': world, 'steps': steps, 'warmup': max(warmup, 6), 'ms_per_step': ms / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32 (TF32 convolutions, fp32 GEMMs)' + ('' if vdt is None else ', bf16 value storage'),
        'data': 'synthetic',
        'config': {'workload': 'pavenet_step', 'description': 'PAVE-Net R-50, T=3 frames at 800x1333, '
                   '1 clip per GPU, forward + backward + gradient all-reduce (NCCL) + grad-clip + AdamW',
                   'trainable_params': n_params, 'parallelism': 'clip-sharded data parallel x%d' % world,
                   'cuda_graphs': bool(args.graphs), 'grad_exchange': args.grad_exchange,
                   'l2_policy': 'inputs larger than L2 (activations of a 3x800x1333 clip)'},
        'collective': {'backend': 'nccl' if world > 1 else None, 'ranks': world,
                       'bytes_all_reduced_per_step': 4 * n_params if world > 1 else 0,
                       'exposed_all_reduce_ms': exposed_ms,
                       'buckets': None if flat is None else [b - a for a, b in flat.ranges]},
        # kernels of this repository's library: launched directly + replayed from CUDA graphs
        'clocks': clock_info,
        'gpu_launches': int(_capi.launch_count() + model.library_launches_replayed() - launches0),
        'final_loss': float(loss)}
    del model, opt, flat, ddp, batches
    torch.cuda.empty_cache()
    return rec


def run_model_step(args, world, rank, local):
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rec = measure_model_step(args, world, rank, local, args.steps, args.warmup)
    rec.update({'roofline': None, 'cpu_baseline': None, 'e2e': None})
    if rank == 0:
        emit(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def measure_gpu_baseline(probs, bufs, dims, q_per_step, steps):
    """The reference's own CUDA kernels (ms_deform_attn_cuda_kernel.cuh recompiled for sm_100a,
    oracle/_ref/libmsda_refcuda.so) on the same device tensors, launched as the reference's host
    code launches them; the zero-fills its Python wrapper performs (output, three gradients,
    multi_scale_deform_attn.py:72-74, ms_deform_attn_cuda.cu:247) are timed as their own item."""
    from oracle import msda_oracle as O
    if not O.refcuda_available():
        return {'unavailable': 'oracle/_ref/libmsda_refcuda.so not built (make -C oracle refcuda, '
                               'needs the reference checkout)'}
    outs = [torch.empty((dims['B'], dims['Q'], dims['M'] * dims['D']), device=p['value'].device)
            for p in probs]

    def step(i, ev=None):
        p, b, o = probs[i % len(probs)], bufs[i % len(probs)], outs[i % len(probs)]
        if ev:
            ev[0].record()
        O.refcuda_forward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], out=o)
        if ev:
            ev[1].record()
        b['grad_value'].zero_()
        b['grad_loc'].zero_()
        b['grad_aw'].zero_()
        if ev:
            ev[2].record()
        O.refcuda_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], p['grad_out'],
                           b['grad_value'], b['grad_loc'], b['grad_aw'])
        if ev:
            ev[3].record()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    for i in range(steps):
        step(i, evs[i])
    torch.cuda.synchronize()
    fwd = sum(e[0].elapsed_time(e[1]) for e in evs) / steps
    zero = sum(e[1].elapsed_time(e[2]) for e in evs) / steps
    bwd = sum(e[2].elapsed_time(e[3]) for e in evs) / steps
    total = evs[0][0].elapsed_time(evs[-1][3]) / steps
    return {'kind': 'reference CUDA kernels (ms_deform_attn_cuda_kernel.cuh) recompiled for sm_100a, '
                    'same tensors, same GPU',
            'kernel_ms': {'fwd': fwd, 'grad_zero_fill': zero, 'bwd': bwd},
            'ms_per_step': total, 'value': q_per_step / (total * 1e-3), 'unit': 'queries/s',
            'steps': steps}


def load_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:  # noqa: BLE001
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def load_traffic(workload, kernel_key):
    """dram bytes per launch of this workload's kernel from the committed
    `ncu --set full` capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(path) as f:
            return json.load(f).get(workload, {}).get(kernel_key)
    except Exception:  # noqa: BLE001
        return None


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='encoder_cfg2',
                    choices=sorted(WORKLOADS) + list(MODEL_WORKLOADS))
    ap.add_argument('--value-dtype', default='f32', choices=['f32', 'bf16'])
    ap.add_argument('--grad-exchange', default='flat', choices=['flat', 'overlap', 'ddp'],
                    help='pavenet_step: one all-reduce of the flat gradient buffer after the backward (default: '
                         'measured fastest at N=8, profiles/r02_multi_gpu.txt), stage buckets all-reduced '
                         'underneath the backward, or torch DDP')
    ap.add_argument('--model-steps', type=int, default=-1,
                    help='op workloads: also time this many steps of the clip-sharded PAVE-Net R-50 training step '
                         '(BASELINE config 4, NCCL gradient all-reduce) and append it as `pavenet_step`; '
                         '-1 = 10 for the default workload, 0 for the others; 0 = skip')
    ap.add_argument('--graphs', type=int, default=1, help='pavenet_step: run backbone, encoder and pose decoder as CUDA graphs (fwd + bwd)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-gpu-baseline', action='store_true',
                    help='skip timing the reference CUDA kernels (oracle/_ref/libmsda_refcuda.so)')
    ap.add_argument('--sets', type=int, default=4, help='distinct input sets rotated per step')
    ap.add_argument('--frames', type=int, default=0, help='override the batch entries (frames) per step of an op workload: the batch sweep of BASELINE config 5')
    ap.add_argument('--fused', action='store_true', help='time the fused-prologue kernels (offsets/logits in, softmax + location transform in-kernel) on the same problem')
    ap.add_argument('--piece-mb', type=float, default=0, help='e2e: upload MiB per pipeline piece (0 = library default)')
    ap.add_argument('--e2e-depth', type=int, default=3, help='e2e, queued form: host-buffer calls in flight (workspaces)')
    ap.add_argument('--fold-clear', default='auto', choices=['auto', '0', '1'],
                    help='zero-fill grad_value inside the forward call (msda_forward_clear) instead of a '
                         'separate memset between forward and backward; auto = for the small-Q (pose) workloads, '
                         'whose persistent forward kernel folds the fill in')
    ap.add_argument('--launch', default='auto', choices=['auto', 'eager', 'graph'],
                    help='how the timed steps are launched: eager = one Python / C-ABI call per kernel with CUDA '
                         'events around every kernel; graph = the same calls captured once into a CUDA graph '
                         '(`sets` steps per graph) and replayed, which is what a launch-bound small-Q caller does '
                         '(tools/exp_launch_bound.py: the host needs 45-57 us per step, PETR\'s kernels 39); '
                         'auto = graph for the small-Q (pose) workloads.  The eager leg always runs: it provides '
                         'the per-kernel times')
    ap.add_argument('--option', action='append', default=[], metavar='NAME=INT',
                    help='library kernel-selection knob (msda_set_option), e.g. flat=0, bwd_variant=2')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world, rank, local = dist_setup(args.gpus)
    if args.workload in MODEL_WORKLOADS:
        if args.impl == 'reference':
            if rank == 0:
                emit(json.dumps({'impl': 'reference', 'unavailable':
                                  'the reference model cannot be imported here (mmcv/mmdet/opera '
                                  'dependencies absent); only the op has a CPU reference arm'}))
            return
        run_model_step(args, world, rank, local)
        return
    if args.impl == 'reference':
        run_reference_arm(args, world, rank)
        return

    import torch.distributed as dist
    import pavenet_b200
    from pavenet_b200 import _capi, clip_sharding
    from pavenet_b200.functional import (MultiScaleDeformableAttnFunction, ms_deform_attn_backward,
                                         ms_deform_attn_forward)

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU path to time); '
                         'use --impl reference for the CPU baseline')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    _capi.load()
    for opt in args.option:
        name, val = opt.split('=')
        _capi.set_option(name, int(val))

    wl = args.workload
    cfg = WORKLOADS[wl]
    fold_clear = (cfg['kind'] == 'pose') if args.fold_clear == 'auto' else args.fold_clear == '1'
    vdt = torch.float32 if args.value_dtype == 'f32' else torch.bfloat16
    # every rank owns its own clips: weak scaling, `sets` distinct clips per rank
    probs = [make_problem(wl, seed=1000 * rank + i, device=device, value_dtype=vdt,
                          frames=args.frames or None)
             for i in range(args.sets)]
    dims = probs[0]['dims']
    q_per_step = dims['B'] * dims['Q']
    if args.fused:
        # the same samples expressed as the modules' raw projections:
        # loc = ref + off / (W_l, H_l) with ref = 0.5, aw = softmax(logits)
        lib = _capi.load()
        norm = torch.stack([probs[0]['shapes'][:, 1], probs[0]['shapes'][:, 0]], -1).float()
        for p in probs:
            B_, Q_, M_, L_, P_ = p['aw'].shape
            p['ref'] = torch.full((B_, Q_, L_, 1, 2), 0.5, device=device)
            p['off'] = ((p['loc'] - 0.5) * norm[None, None, None, :, None, :]).contiguous()
            p['logit'] = p['aw'].clamp_min(1e-30).log().reshape(B_, Q_, M_, L_ * P_).contiguous()
            p['stats'] = torch.empty((B_, Q_, M_, 2), device=device)
            p['out'] = torch.empty((B_, Q_, M_ * dims['D']), device=device)
    gv_dtype = torch.float32
    bufs = [dict(grad_value=torch.empty(p['value'].shape, dtype=gv_dtype, device=device),
                 grad_loc=torch.empty_like(p['loc']), grad_aw=torch.empty_like(p['aw']))
            for p in probs]
    footprint = sum(t.numel() * t.element_size() for p in probs for t in p.values()
                    if isinstance(t, torch.Tensor))
    footprint += sum(t.numel() * t.element_size() for b in bufs for t in b.values())

    vcode = 0 if vdt == torch.float32 else 2

    def step(i, ev=None):
        p, b = probs[i % args.sets], bufs[i % args.sets]
        stream = torch.cuda.current_stream().cuda_stream
        if ev:
            ev[0].record()
        if args.fused:
            _capi.check(lib.msda_fused_forward(
                p['value'].data_ptr(), p['shapes'].data_ptr(), p['lsi'].data_ptr(),
                p['off'].data_ptr(), p['logit'].data_ptr(), p['ref'].data_ptr(), None,
                p['out'].data_ptr(), p['stats'].data_ptr(), dims['B'], dims['S'], dims['M'],
                dims['D'], dims['L'], dims['Q'], dims['P'], 1, vcode,
                b['grad_value'].data_ptr() if fold_clear else None,
                b['grad_value'].numel() * b['grad_value'].element_size() if fold_clear else 0,
                stream), 'msda_fused_forward')
            out = p['out']
        else:
            out = ms_deform_attn_forward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'], 64,
                                         clear=b['grad_value'] if fold_clear else None)
        if ev:
            ev[1].record()
        if not fold_clear:
            b['grad_value'].zero_()
        if ev:
            ev[2].record()
        if args.fused:
            _capi.check(lib.msda_fused_backward(
                p['value'].data_ptr(), p['shapes'].data_ptr(), p['lsi'].data_ptr(),
                p['off'].data_ptr(), p['logit'].data_ptr(), p['ref'].data_ptr(), None,
                p['stats'].data_ptr(), p['out'].data_ptr(), p['grad_out'].data_ptr(),
                b['grad_value'].data_ptr(),
                b['grad_loc'].data_ptr(), b['grad_aw'].data_ptr(), None, dims['B'], dims['S'],
                dims['M'], dims['D'], dims['L'], dims['Q'], dims['P'], 1, vcode, stream),
                'msda_fused_backward')
        else:
            ms_deform_attn_backward(p['value'], p['shapes'], p['lsi'], p['loc'], p['aw'],
                                    p['grad_out'], b['grad_value'], b['grad_loc'], b['grad_aw'], 64)
        if ev:
            ev[3].record()
        return out

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = _capi.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        step(i, evs[i])
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clock_info = clocks.stop()
    launches = _capi.launch_count() - launches0

    elapsed_ms = t_start.elapsed_time(t_end)
    eager_elapsed_ms = elapsed_ms
    launch_mode = 'eager'
    use_graph = args.launch == 'graph' or (args.launch == 'auto' and cfg['kind'] == 'pose')
    if use_graph:
        # second timed leg: the same K steps, captured `sets` at a time into one CUDA graph and replayed
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(args.sets):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = _capi.launch_count()
        with torch.cuda.graph(graph):
            for i in range(args.sets):
                step(i)
        per_graph = _capi.launch_count() - n0
        reps, rest = divmod(args.steps, args.sets)
        for _ in range(max(3, -(-args.warmup // args.sets))):
            graph.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = ClockSampler(local)
        clocks.start()
        n0 = _capi.launch_count()
        t_start.record()
        for _ in range(reps):
            graph.replay()
        for i in range(rest):
            step(i)
        t_end.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clock_info = clocks.stop()
        launches = reps * per_graph + (_capi.launch_count() - n0)   # replayed + directly launched
        elapsed_ms = t_start.elapsed_time(t_end)
        launch_mode = 'cuda_graph (%d steps per graph, %d replays + %d eager steps)' % (args.sets, reps, rest)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    zero_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    bwd_ms = sum(e[2].elapsed_time(e[3]) for e in evs) / args.steps
    elapsed_max = clip_sharding.max_over_ranks(elapsed_ms, device)
    total_q = clip_sharding.sum_over_ranks(q_per_step * args.steps, device)
    value = total_q / (elapsed_max * 1e-3)

    # ---- end to end with HOST buffers: H2D of the inputs, kernels, D2H of all results ----
    e2e = None
    e2e_autograd = None
    if not args.no_e2e:
        # pinned buffers on the NUMA node of this rank's GPU (matters when N ranks share a host)
        host_affinity = bind_to_gpu_numa(local) if world > 1 else {'bound': False}
        p = probs[0]
        host = {k: torch.empty(p[k].shape, dtype=p[k].dtype).pin_memory()
                for k in ('value', 'loc', 'aw', 'grad_out')}
        for k in host:
            host[k].copy_(p[k])
        shapes_h, lsi_h = p['shapes'].cpu(), p['lsi'].cpu()
        out_h = torch.empty((dims['B'], dims['Q'], dims['M'] * dims['D'])).pin_memory()
        gv_h = torch.empty(p['value'].shape, dtype=torch.float32).pin_memory()
        gl_h = torch.empty(p['loc'].shape).pin_memory()
        ga_h = torch.empty(p['aw'].shape).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host.values())
        d2h = sum(t.numel() * t.element_size() for t in (out_h, gv_h, gl_h, ga_h))
        n_e2e = max(5, min(args.steps, 30))

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                fn()
            e1.record()
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
            # the staged C-ABI call runs on the library's own streams and returns when the
            # results are in host memory: wall clock around blocking calls is the honest
            # timer there; CUDA events on torch's stream cover the autograd variant
            ms = max(wall_ms, e0.elapsed_time(e1))
            ms = clip_sharding.max_over_ranks(ms, device)
            tq = clip_sharding.sum_over_ranks(q_per_step * n_e2e, device)
            return {'value': tq / (ms * 1e-3), 'unit': 'queries/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms / n_e2e, 'steps': n_e2e}

        # (1) the C-ABI host-buffer call a non-PyTorch caller binds (pipelined inside the library)
        hws = pavenet_b200.HostWorkspace()
        if args.piece_mb > 0:
            hws.set_piece_bytes(int(args.piece_mb * (1 << 20)))

        def e2e_capi():
            hws.forward_backward(host['value'], shapes_h, lsi_h, host['loc'], host['aw'],
                                 host['grad_out'], out=out_h, grad_value=gv_h,
                                 grad_sampling_loc=gl_h, grad_attn_weight=ga_h)

        e2e_blocking = timed(e2e_capi)
        e2e_blocking['api'] = ('msda_forward_backward_host (C ABI, pinned host buffers; upload / kernels / '
                               'download pipelined over batch entries x query chunks), one blocking call per step')

        # (1b) the same entry point, queued: two calls in flight on two workspaces, each with its own
        # result buffers.  Every step still uploads all its inputs and downloads all its results inside
        # the timed region; the next step's first upload runs under this step's last download, which a
        # blocking call leaves idle (0.9 ms of its 6.4, profiles/r02_e2e_link_analysis.txt)
        # three calls in flight, each in the monolithic form (piece size >= the call: every tensor is one copy,
        # 34-68 MB, and the kernels run once over the whole batch) -- 5.14 ms per step against 5.39 for two calls
        # with one piece per batch entry (profiles/r02_e2e_link_analysis.txt)
        slots = [(hws, out_h, gv_h, gl_h, ga_h)]
        for _ in range(max(2, args.e2e_depth) - 1):
            slots.append((pavenet_b200.HostWorkspace(), torch.empty_like(out_h).pin_memory(),
                          torch.empty_like(gv_h).pin_memory(), torch.empty_like(gl_h).pin_memory(),
                          torch.empty_like(ga_h).pin_memory()))
        for s_ in slots:     # (the blocking leg above has already run on slots[0]'s workspace)
            s_[0].set_piece_bytes(int(args.piece_mb * (1 << 20)) if args.piece_mb > 0 else 1 << 40)
        state = {'i': 0}

        def e2e_queued():
            ws_, o_, gv_, gl_, ga_ = slots[state['i'] % len(slots)]
            state['i'] += 1
            ws_.wait()                  # the call queued on this workspace two steps ago
            ws_.forward_backward(host['value'], shapes_h, lsi_h, host['loc'], host['aw'],
                                 host['grad_out'], out=o_, grad_value=gv_, grad_sampling_loc=gl_,
                                 grad_attn_weight=ga_, wait=False)

        def timed_queued():
            for _ in range(2 * len(slots)):
                e2e_queued()
            for s_ in slots:
                s_[0].wait()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_queued()
            for s_ in slots:
                s_[0].wait()            # every result of every step is in host memory
            wall_ms = (time.perf_counter() - t0) * 1e3
            ms = clip_sharding.max_over_ranks(wall_ms, device)
            tq = clip_sharding.sum_over_ranks(q_per_step * n_e2e, device)
            return {'value': tq / (ms * 1e-3), 'unit': 'queries/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms / n_e2e, 'steps': n_e2e}

        e2e_q = timed_queued()
        e2e_q['api'] = ('msda_forward_backward_host_async + msda_workspace_wait (C ABI, pinned host buffers): '
                        '%d calls in flight on as many workspaces with separate result buffers; every step uploads ' % len(slots) +
                        'all inputs and downloads output + all gradients; each call moves every tensor as one copy '
                        '(monolithic form), the overlap is between calls')
        # both are the public entry point; a caller picks the form that suits its host.  With one GPU per
        # host link the queued form wins (5.6 against 6.4 ms); when several ranks share one host memory
        # system (N > 1 on this box) the extra concurrency costs more than the hidden head / tail saves.
        # The two times are max-over-ranks, so every rank takes the same branch.
        if e2e_q['ms_per_step'] <= e2e_blocking['ms_per_step']:
            e2e = dict(e2e_q, mode='queued', blocking=e2e_blocking)
        else:
            e2e = dict(e2e_blocking, mode='blocking', queued=e2e_q)
        e2e['host_affinity'] = host_affinity
        # the last queued results must equal the blocking call's
        e2e['max_abs_diff_vs_blocking'] = float(max((slots[1][k] - slots[0][k]).abs().max() for k in (1, 3, 4)))
        for s_ in slots:
            s_[0].close()

        # (2) the autograd Function PyTorch callers use, copies issued around it on one stream
        gvb_h = torch.empty(p['value'].shape, dtype=p['value'].dtype).pin_memory()

        def e2e_torch():
            v = host['value'].to(device, non_blocking=True).requires_grad_()
            loc = host['loc'].to(device, non_blocking=True).requires_grad_()
            aw = host['aw'].to(device, non_blocking=True).requires_grad_()
            go = host['grad_out'].to(device, non_blocking=True)
            out = MultiScaleDeformableAttnFunction.apply(v, p['shapes'], p['lsi'], loc, aw, 64)
            out.backward(go)
            out_h.copy_(out.detach(), non_blocking=True)
            gvb_h.copy_(v.grad, non_blocking=True)
            gl_h.copy_(loc.grad, non_blocking=True)
            ga_h.copy_(aw.grad, non_blocking=True)

        e2e_autograd = timed(e2e_torch)
        e2e_autograd['api'] = ('MultiScaleDeformableAttnFunction.apply + backward, pinned host '
                               'tensors, one stream (no overlap)')

    # ---- the reference's own CUDA kernels on the same tensors (oracle/_ref, bench infrastructure) ----
    gpu_baseline = None
    if rank == 0 and not args.no_gpu_baseline and vdt == torch.float32:
        gpu_baseline = measure_gpu_baseline(probs, bufs, dims, q_per_step, min(args.steps, 50))

    # ---- BASELINE config 4 in the same run: the clip-sharded training step, NCCL all-reduce on ----
    model_step = None
    n_model = args.model_steps if args.model_steps >= 0 else (10 if wl == 'encoder_cfg2' and not args.fused else 0)
    if n_model > 0:
        try:
            model_step = measure_model_step(args, world, rank, local, n_model, 6)
        except Exception as exc:  # noqa: BLE001 - the op line must still be printed
            model_step = {'error': '%s: %s' % (type(exc).__name__, exc)}

    if rank == 0:
        peak, peak_src = load_peak()
        vb = 4 if vdt == torch.float32 else 2
        traffic_wl = wl if vdt == torch.float32 else wl + '_bf16'
        ab = algorithmic_bytes(dims, value_bytes=vb, grad_value_bytes=4)

        def roof(nbytes, ms, key):
            gbs = nbytes / (ms * 1e-3) / 1e9
            return {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s',
                    'frac': gbs / peak, 'traffic': load_traffic(traffic_wl, key), 'kernel': key,
                    'algorithmic_bytes': nbytes, 'kernel_ms': ms, 'peak_source': peak_src,
                    'frac_of_8TBs_nominal': gbs / 8000.0}

        # Secondary, on-chip rooflines (SURVEY.md section 8d asks for the L1 / L2 limiter next to HBM):
        # the op gathers / reduces 4*L*P value rows per output row, so what bounds the kernels is
        # what an SM can move per clock, not HBM.  Forward peak: the L1 data path, one 128-byte
        # wavefront (= one fp32 value row) per clock and SM at the SM clock sampled during the run
        # (hardware figure; rows that miss L1 are further limited by the 64 B/clk/SM fill path from
        # L2, measured at 1.84 clocks per row, profiles/r01_microbench_gather_rows.txt).  Backward
        # peak: the rate at which the SMs can issue 128-byte reductions into L2, measured
        # (red.global.add.v4.f32 54.0 G rows/s = 5.4 clocks per row and SM; the TMA path
        # cp.reduce.async.bulk reaches the same 49-50 G rows/s, profiles/r02_microbench_tma_reduce.txt).
        corner_rows = 4.0 * dims['B'] * dims['Q'] * dims['M'] * dims['L'] * dims['P']
        n_sms = torch.cuda.get_device_properties(device).multi_processor_count

        def onchip(ms, peak_rows, what, source, **extra):
            rate = corner_rows / (ms * 1e-3) / 1e9
            out = {'bound': what, 'achieved': rate, 'peak': peak_rows, 'unit': 'G rows/s (128-byte value rows)',
                   'frac': rate / peak_rows, 'rows_per_launch': corner_rows,
                   'peak_source': source}
            out.update(extra)
            return out

        kname = _capi.kernel_name(dims['D'], 0, 0 if vdt == torch.float32 else 2)
        line = {
            'metric': 'deform-attn fwd+bwd queries/s', 'value': value, 'unit': 'queries/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': elapsed_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32' if vdt == torch.float32 else 'f32 (bf16 value storage)',
            'data': 'synthetic',
            'config': workload_config(wl, dims, args.sets, world),
            'kernel': kname + (' + fused prologue' if args.fused else ''),
            'footprint_gb_per_gpu': footprint / 1e9,
            'roofline': roof(ab['bwd'], bwd_ms, 'msda_bwd_rows_kernel'),
            'roofline_fwd': roof(ab['fwd'], fwd_ms, 'msda_fwd_rows_kernel'),
            'roofline_step': roof(ab['fwd'] + ab['bwd'],
                                  elapsed_max / args.steps if use_graph else fwd_ms + zero_ms + bwd_ms, 'step'),
            'launch': launch_mode,
            'ms_per_step_eager': eager_elapsed_ms / args.steps,
            'roofline_onchip': None if vdt != torch.float32 else {   # the peaks are fp32-row figures
                'bwd': onchip(bwd_ms, 53.97, 'sm_to_l2_reduction',
                              'micro-benchmark, profiles/r01_microbench_scatter_rows.txt'),
                'fwd': onchip(fwd_ms, n_sms * (clock_info.get('sm_mhz') or 1965) * 1e-3, 'l1_data_path',
                              'hardware: 128 B per clock and SM x %d SMs x sampled SM clock' % n_sms,
                              measured_random_row_gather={'l2_resident': 157.63, 'l1_resident': 259.55,
                                                          'source': 'profiles/r01_microbench_gather_rows.txt'})},
            'kernel_ms': {'fwd': fwd_ms, 'grad_value_zero_fill': zero_ms, 'bwd': bwd_ms,
                          'grad_value_zero_fill_folded_into_fwd': bool(fold_clear)},
            'kernel_families': {k: v for k, v in _capi.family_counts().items() if v},
            'options': args.option,
            'clocks': clock_info, 'gpu_launches': int(launches), 'e2e': e2e,
            'e2e_autograd': e2e_autograd,
            'gpu_baseline': gpu_baseline,
            'pavenet_step': model_step,
        }
        if gpu_baseline:
            gpu_baseline['speedup'] = {
                'fwd': gpu_baseline['kernel_ms']['fwd'] / fwd_ms,
                'bwd': gpu_baseline['kernel_ms']['bwd'] / bwd_ms,
                'step': gpu_baseline['ms_per_step'] / (fwd_ms + zero_ms + bwd_ms)}
        if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N = 1 only
            line['cpu_baseline'] = measure_cpu_baseline(wl)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
