"""GPU parity of every kernel FAMILY behind the dispatcher, each forced through
`msda_set_option` and checked against the C oracle, with the library's per-family launch
counters asserting that the family under test is the one that ran (a silent fallback to
another kernel would otherwise pass).

Families (include/pavenet_msda.h, msda_launch_count_family): rows (large Q), flat (small Q:
persistent grid over flattened (row, 32-sample chunk) space), generic (any D / fp64), each
forward + backward, plain and fused-prologue.
Tolerances: fp32 outputs <= 1e-4, gradients <= 1e-3 (max|a-b| / max|b|), as BASELINE.json.
"""
import pytest
import torch

from conftest import rel_err
from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

MID = [(28, 40), (14, 20), (7, 10), (4, 5)]
FLAT_ORDER_DEFAULT = 1   # library default of the flat kernels' piece order (msda_kernels.h)


@pytest.fixture()
def lib_options():
    """Set kernel-selection knobs for one test and restore the defaults afterwards."""
    from pavenet_b200 import _capi
    touched = []

    def set_(name, value):
        touched.append(name)
        _capi.set_option(name, value)

    yield set_
    defaults = {'flat': 1, 'force_generic': 0, 'fwd_split': 0, 'bwd_split': 0,
                'bwd_variant': 0, 'fwd_variant': 0, 'clear_mode': 0, 'flat_order': FLAT_ORDER_DEFAULT, 'clear_policy': 2}
    for name in touched:
        _capi.set_option(name, defaults[name])


def _problem(seed, B, Q, M, D, P, shapes, spread=0.15):
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.tensor(shapes, dtype=torch.long)
    L = len(shapes)
    S = int(shapes_t.prod(1).sum())
    value = torch.randn(B, S, M, D, generator=g)
    loc = torch.rand(B, Q, M, L, P, 2, generator=g) * (1 + 2 * spread) - spread
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P)
    go = torch.randn(B, Q, M * D, generator=g)
    return value, shapes_t, O.level_start_index(shapes_t), loc, aw, go


def _run(value, shapes, lsi, loc, aw, go, value_dtype=torch.float32):
    import pavenet_b200
    fn = pavenet_b200.MultiScaleDeformableAttnFunction.apply
    v = value.cuda().to(value_dtype).requires_grad_()
    l = loc.cuda().requires_grad_()
    a = aw.cuda().requires_grad_()
    out = fn(v, shapes.cuda(), lsi.cuda(), l, a, 64)
    out.backward(go.cuda())
    return out.detach(), v.grad, l.grad, a.grad


def _delta(before, after):
    return {k: after[k] - before[k] for k in after if after[k] != before[k]}


FLAT_SHAPES = [
    # B, Q, M, D, P, levels
    (1, 300, 8, 32, 17, MID),            # PETR pose attention: L*P = 68 (2 chunks + 4 samples)
    (1, 60, 8, 32, 17, MID * 5),         # T=5 fused pose decoder: L*P = 340
    (2, 37, 8, 32, 15, MID * 3),         # T=3: L*P = 180
    (1, 5, 2, 32, 3, MID),               # L*P = 12: one partial chunk per row
    (2, 9, 3, 32, 8, MID),               # L*P = 32: exactly one chunk
    (1, 2500, 8, 32, 4, MID),            # many rows, 16 samples: more chunks than warps
    (1, 40, 4, 16, 9, MID),              # D = 16
    (1, 33, 2, 64, 11, MID[1:]),         # D = 64
    (3, 1, 1, 32, 1, [(1, 1)]),          # degenerate
]


@pytest.mark.parametrize('order', [0, 1])
@pytest.mark.parametrize('B,Q,M,D,P,shapes', FLAT_SHAPES)
def test_flat_kernels_match_oracle(lib_options, B, Q, M, D, P, shapes, order):
    """order 1: every warp walks its piece of the (row, chunk) space from chunk 0 upwards."""
    from pavenet_b200 import _capi
    lib_options('flat', 2)
    lib_options('flat_order', order)
    prob = _problem(7 * Q + D, B, Q, M, D, P, shapes)
    before = _capi.family_counts()
    out, gv, gl, ga = _run(*prob)
    ran = _delta(before, _capi.family_counts())
    assert ran == {'fwd_flat': 1, 'bwd_flat': 1}, ran
    value, shapes_t, lsi, loc, aw, go = prob
    ref = O.c_forward(value, shapes_t, lsi, loc, aw)
    rgv, rgl, rga = O.c_backward(value, shapes_t, lsi, loc, aw, go)
    assert rel_err(out, ref) < 1e-4 and rel_err(out, ref) < 5e-6
    assert rel_err(gv, rgv) < 1e-3
    assert rel_err(gl, rgl) < 1e-3 and rel_err(gl, rgl) < 5e-5
    assert rel_err(ga, rga) < 1e-3


def test_flat_kernels_randomised_shapes(lib_options):
    """30 seeded random problems forced through the flat family: D in {16, 32, 64}, ragged level
    sizes incl. 1xN maps, 1..20 levels (frames-as-levels), 1..19 points, pieces that start and end
    anywhere inside a row (both piece orders), locations far outside the map and a few non-finite
    ones (they must contribute nothing, as in the reference) -- forward and all three gradients
    against the C oracle, and the folded zero-fill checked on the way."""
    import random
    from pavenet_b200 import _capi
    from pavenet_b200.functional import ms_deform_attn_backward, ms_deform_attn_forward
    rng = random.Random(20261017)
    lib_options('flat', 2)
    for trial in range(30):
        D = rng.choice([16, 32, 32, 64])
        M = rng.choice([1, 2, 4, 8])
        L = rng.choice([1, 2, 4, 5, 12, 20])
        P = rng.randint(1, 19)
        B = rng.randint(1, 3)
        Q = rng.choice([1, 3, 17, 64, 300, 700])
        shapes = [(rng.randint(1, 20), rng.randint(1, 20)) for _ in range(L)]
        lib_options('flat_order', trial % 2)
        value, shapes_t, lsi, loc, aw, go = _problem(5000 + trial, B, Q, M, D, P, shapes,
                                                      spread=rng.choice([0.0, 0.15, 0.8]))
        if trial % 3 == 0:      # non-finite locations: the range test fails, the sample is skipped
            flat_loc = loc.view(-1)
            idx = torch.randint(0, flat_loc.numel(), (min(7, flat_loc.numel()),),
                                generator=torch.Generator().manual_seed(trial))
            flat_loc[idx] = torch.tensor([float('nan'), float('inf'), -float('inf')])[idx % 3]
        before = _capi.family_counts()
        args = (value.cuda(), shapes_t.cuda(), lsi.cuda(), loc.cuda(), aw.cuda())
        gv = torch.full(value.shape, 7.0, device='cuda')
        out = ms_deform_attn_forward(*args, 64, clear=gv)
        assert float(gv.abs().max()) == 0.0
        gl, ga = torch.empty_like(args[3]), torch.empty_like(args[4])
        ms_deform_attn_backward(*args, go.cuda(), gv, gl, ga, 64)
        ran = _delta(before, _capi.family_counts())
        assert ran == {'fwd_flat': 1, 'bwd_flat': 1}, (trial, ran)
        ref = O.c_forward(value, shapes_t, lsi, loc, aw)
        rgv, rgl, rga = O.c_backward(value, shapes_t, lsi, loc, aw, go)
        tag = (trial, B, Q, M, D, L, P, shapes)
        assert torch.isfinite(out).all() and torch.isfinite(gv).all(), tag
        assert rel_err(out, ref) < 1e-4, tag
        assert rel_err(gv, rgv) < 1e-3, tag
        assert rel_err(gl, rgl) < 1e-3, tag
        assert rel_err(ga, rga) < 1e-3, tag


@pytest.mark.parametrize('B,Q,M,D,P,shapes', FLAT_SHAPES[:4] + FLAT_SHAPES[6:8])
def test_rows_kernels_forced_on_the_same_shapes(lib_options, B, Q, M, D, P, shapes):
    from pavenet_b200 import _capi
    lib_options('flat', 0)
    prob = _problem(7 * Q + D, B, Q, M, D, P, shapes)
    before = _capi.family_counts()
    out, gv, gl, ga = _run(*prob)
    ran = _delta(before, _capi.family_counts())
    assert ran == {'fwd_rows': 1, 'bwd_rows': 1}, ran
    value, shapes_t, lsi, loc, aw, go = prob
    assert rel_err(out, O.c_forward(value, shapes_t, lsi, loc, aw)) < 5e-6
    rgv, rgl, rga = O.c_backward(value, shapes_t, lsi, loc, aw, go)
    assert rel_err(gv, rgv) < 1e-3 and rel_err(gl, rgl) < 5e-5 and rel_err(ga, rga) < 1e-3


@pytest.mark.parametrize('D,Q,P', [(32, 300, 17), (16, 120, 9), (64, 50, 11)])
def test_flat_kernels_bf16_value(lib_options, D, Q, P):
    """bf16 value storage through the flat family: against the oracle run on the same
    rounded value (tight), fp32 gradient accumulation."""
    from pavenet_b200 import _capi
    lib_options('flat', 2)
    value, shapes_t, lsi, loc, aw, go = _problem(11 + D, 1, Q, 4, D, P, MID)
    v16 = value.to(torch.bfloat16)
    before = _capi.family_counts()
    out, gv, gl, ga = _run(v16.float(), shapes_t, lsi, loc, aw, go, value_dtype=torch.bfloat16)
    assert _delta(before, _capi.family_counts()) == {'fwd_flat': 1, 'bwd_flat': 1}
    ref = O.c_forward(v16.float(), shapes_t, lsi, loc, aw)
    rgv, rgl, rga = O.c_backward(v16.float(), shapes_t, lsi, loc, aw, go)
    assert rel_err(out, ref) < 2e-5
    assert rel_err(gl, rgl) < 1e-3 and rel_err(ga, rga) < 1e-3
    assert gv.dtype == torch.bfloat16 and rel_err(gv.float(), rgv) < 4e-3   # one rounding to bf16


def test_forward_clear_zero_fills_the_buffer(lib_options):
    """msda_forward_clear: the buffer handed to the forward comes back all zeros, for the flat
    family (fill folded into the kernel) and the rows family (memset), any size / tail."""
    from pavenet_b200.functional import ms_deform_attn_forward
    value, shapes_t, lsi, loc, aw, _ = _problem(3, 1, 60, 8, 32, 17, MID * 2)
    args = (value.cuda(), shapes_t.cuda(), lsi.cuda(), loc.cuda(), aw.cuda())
    ref = O.c_forward(value, shapes_t, lsi, loc, aw)
    # clear_mode 0: streaming stores between the chunks of the flat kernel; 2: TMA bulk stores of a
    # zeroed shared-memory tile queued at kernel start; 1: memset on a side stream
    # clear_policy: cache policy of the streaming-store form (2 = L2 evict-last, the default)
    for flat, mode, policy in ((2, 0, 2), (2, 0, 0), (2, 0, 1), (2, 2, 2), (2, 1, 2), (0, 0, 2), (0, 2, 2)):
        lib_options('flat', flat)
        lib_options('clear_mode', mode)
        lib_options('clear_policy', policy)
        for n in (4, 1024, 4 * 1000 * 1000 + 4, 12345 * 4, 7, 1001):      # incl. sizes that are not 16-byte multiples
            buf = torch.full((n,), 3.0, device='cuda')
            guard = torch.full((64,), 5.0, device='cuda')
            out = ms_deform_attn_forward(*args, 64, clear=buf)
            assert float(buf.abs().max()) == 0.0, (flat, mode, policy, n)
            assert float(guard.min()) == 5.0
            assert rel_err(out, ref) < 5e-6


def test_function_reuses_the_cleared_buffer_once():
    """The autograd Function zero-fills grad_value during the forward call; a second backward
    through a retained graph must not reuse the dirty buffer."""
    import pavenet_b200
    fn = pavenet_b200.MultiScaleDeformableAttnFunction.apply
    value, shapes_t, lsi, loc, aw, go = _problem(5, 1, 50, 8, 32, 17, MID)
    v = value.cuda().requires_grad_()
    out = fn(v, shapes_t.cuda(), lsi.cuda(), loc.cuda(), aw.cuda(), 64)
    out.backward(go.cuda(), retain_graph=True)
    g1 = v.grad.clone()
    v.grad = None
    out.backward(go.cuda())
    assert rel_err(v.grad, g1) < 1e-6
    rgv, _, _ = O.c_backward(value, shapes_t, lsi, loc, aw, go)
    assert rel_err(g1, rgv) < 1e-3


@pytest.mark.parametrize('flat', [2, 0])
@pytest.mark.parametrize('Q,P,levels,R,with_scale,D', [
    (40, 17, MID * 3, 17, True, 32),      # pose decoder: one reference per point, box scale
    (300, 4, MID, 1, False, 32),          # encoder style: one reference per level, / (W, H)
    (7, 15, MID * 5, 15, True, 32),
    (90, 4, MID, 1, False, 16),           # other head sizes of the fused path
    (33, 9, MID * 2, 9, True, 64),
])
def test_fused_kernels_match_unfused_chain(lib_options, flat, Q, P, levels, R, with_scale, D):
    """softmax + location transform in the kernel (flat and rows families) against the
    op-by-op chain run through autograd on the plain op, all gradients."""
    import pavenet_b200
    from pavenet_b200 import _capi
    from pavenet_b200.functional import FusedMultiScaleDeformableAttnFunction
    lib_options('flat', flat)
    g = torch.Generator().manual_seed(Q + P)
    shapes_t = torch.tensor(levels, dtype=torch.long)
    L, B, M = len(levels), 2, 8
    S = int(shapes_t.prod(1).sum())
    lsi = O.level_start_index(shapes_t)
    dev = 'cuda'
    value = torch.randn(B, S, M, D, generator=g).to(dev)
    off = (torch.randn(B, Q, M, L, P, 2, generator=g) * (0.1 if with_scale else 2.0)).to(dev)
    logit = torch.randn(B, Q, M, L * P, generator=g).to(dev)
    ref = torch.rand(B, Q, L, R, 2, generator=g).to(dev)
    scale = (torch.rand(B, Q, L, 2, generator=g) * 0.5).to(dev) if with_scale else None
    go = torch.randn(B, Q, M * D, generator=g).to(dev)

    def leaves():
        return [t.detach().clone().requires_grad_() if t is not None else None
                for t in (value, off, logit, ref, scale)]

    v, o, lg, r, sc = leaves()
    before = _capi.family_counts()
    out = FusedMultiScaleDeformableAttnFunction.apply(v, shapes_t.to(dev), lsi.to(dev), o, lg, r, sc)
    out.backward(go)
    ran = _delta(before, _capi.family_counts())
    fam = 'flat' if flat == 2 else 'rows'
    assert ran == {'fwd_%s_fused' % fam: 1, 'bwd_%s_fused' % fam: 1}, ran

    v2, o2, lg2, r2, sc2 = leaves()
    w = torch.softmax(lg2, -1).view(B, Q, M, L, P)
    if sc2 is not None:
        loc = r2[:, :, None] + o2 * sc2[:, :, None, :, None, :]
    else:
        norm = torch.stack([shapes_t[:, 1], shapes_t[:, 0]], -1).float().to(dev)
        loc = r2[:, :, None] + o2 / norm[None, None, None, :, None, :]
    out2 = pavenet_b200.MultiScaleDeformableAttnFunction.apply(
        v2, shapes_t.to(dev), lsi.to(dev), loc.contiguous(), w.contiguous(), 64)
    out2.backward(go)
    assert rel_err(out, out2) < 1e-5
    assert rel_err(v.grad, v2.grad) < 2e-4
    assert rel_err(o.grad, o2.grad) < 2e-4
    assert rel_err(lg.grad, lg2.grad) < 2e-4
    assert rel_err(r.grad, r2.grad) < 2e-4
    if sc is not None:
        assert rel_err(sc.grad, sc2.grad) < 2e-4


ENC_LEVELS = [(50, 84), (25, 42), (13, 21), (7, 11)]


@pytest.mark.parametrize('variant', [('fwd_variant', 2), ('bwd_variant', 2), ('bwd_variant', 3)])
def test_large_q_kernel_variants_match_oracle(lib_options, variant):
    """Encoder-shaped problem (queries = pixels, coherent locations so neighbouring queries DO
    collide on value rows) through the alternative large-Q kernels: the head-affine persistent
    forward (fwd_variant 2), the warp-aggregated-atomics backward (bwd_variant 2) and the backward that
    privatises the coarsest level in shared memory (bwd_variant 3, 128-bit CAS accumulation)."""
    from pavenet_b200 import _capi
    name, val = variant
    lib_options(name, val)
    lib_options('flat', 0)
    g = torch.Generator().manual_seed(77)
    shapes_t = torch.tensor(ENC_LEVELS, dtype=torch.long)
    lsi = O.level_start_index(shapes_t)
    S = int(shapes_t.prod(1).sum())
    B, M, D, L, P, Q = 2, 8, 32, 4, 4, S
    ref = []
    for h, w in ENC_LEVELS:
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        ref.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    ref = torch.cat(ref)
    norm = torch.tensor([[w, h] for h, w in ENC_LEVELS], dtype=torch.float32)
    loc = ref[None, :, None, None, None, :] + torch.randn(B, Q, M, L, P, 2, generator=g) * 0.6 / \
        norm[None, None, None, :, None, :]
    value = torch.randn(B, S, M, D, generator=g)
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P)
    go = torch.randn(B, Q, M * D, generator=g)
    before = _capi.family_counts()
    out, gv, gl, ga = _run(value, shapes_t, lsi, loc, aw, go)
    assert _delta(before, _capi.family_counts()) == {'fwd_rows': 1, 'bwd_rows': 1}
    assert rel_err(out, O.c_forward(value, shapes_t, lsi, loc, aw)) < 5e-6
    rgv, rgl, rga = O.c_backward(value, shapes_t, lsi, loc, aw, go)
    assert rel_err(gv, rgv) < 1e-3 and rel_err(gv, rgv) < 2e-5
    assert rel_err(gl, rgl) < 1e-3 and rel_err(ga, rga) < 1e-3
    # elementwise, not only max-normalised: |a-b| <= 1e-3 (|b| + rms(b)) for every element
    b = rgv.double()
    err = (gv.cpu().double() - b).abs() / (b.abs() + float(b.pow(2).mean().sqrt()))
    assert float(err.max()) < 1e-3


def test_wide_load_forward_variants_match_oracle(lib_options):
    """256-bit value loads (ld.global.nc.v8.f32): rows family (fwd_variant 3) on an encoder-like
    problem, D = 32 and 64."""
    from pavenet_b200 import _capi
    lib_options('fwd_variant', 3)
    lib_options('flat', 0)
    for D, Q in ((32, 1111), (64, 700)):
        prob = _problem(100 + D, 2, Q, 4, D, 4, MID)
        before = _capi.family_counts()
        out, gv, gl, ga = _run(*prob)
        assert _delta(before, _capi.family_counts()) == {'fwd_rows': 1, 'bwd_rows': 1}
        value, shapes_t, lsi, loc, aw, go = prob
        assert rel_err(out, O.c_forward(value, shapes_t, lsi, loc, aw)) < 5e-6


@pytest.mark.parametrize('levels', [[(50, 84), (25, 42), (13, 21), (7, 11)], [(37, 53), (19, 27), (10, 14), (5, 7)],
                                    [(64, 64), (32, 32)]])
@pytest.mark.parametrize('variant', [5, 6])
def test_tile_staged_forward_matches_oracle(lib_options, levels, variant):
    """fwd_variant 5: level windows staged in shared memory (msda_fwd_tile.cu); 6: the same
    patch-per-block walk on cached global loads.  Encoder geometry (queries = pixels), coherent
    locations with offsets up to ~12 px so that part of the samples falls outside the staged
    windows (global-memory path) and part outside the maps."""
    from pavenet_b200 import _capi
    lib_options('fwd_variant', variant)
    lib_options('flat', 0)
    g = torch.Generator().manual_seed(len(levels))
    shapes_t = torch.tensor(levels, dtype=torch.long)
    lsi = O.level_start_index(shapes_t)
    S = int(shapes_t.prod(1).sum())
    B, M, D, L, P, Q = 2, 8, 32, len(levels), 4, S
    ref = []
    for h, w in levels:
        ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing='ij')
        ref.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    ref = torch.cat(ref)
    norm = torch.tensor([[w, h] for h, w in levels], dtype=torch.float32)
    off = torch.randn(B, Q, M, L, P, 2, generator=g) * 3.0
    loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    value = torch.randn(B, S, M, D, generator=g)
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P)
    go = torch.randn(B, Q, M * D, generator=g)
    before = _capi.family_counts()
    out, gv, gl, ga = _run(value, shapes_t, lsi, loc, aw, go)
    assert _delta(before, _capi.family_counts()) == {'fwd_tile': 1, 'bwd_rows': 1}
    assert rel_err(out, O.c_forward(value, shapes_t, lsi, loc, aw)) < 5e-6
