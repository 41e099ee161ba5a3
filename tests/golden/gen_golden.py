"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/gen_golden.py [/root/reference]

Nothing here is imported at test time; tests read the committed `.npz` files.
No reference source is copied into this repository: the script parses the
reference files where they lie, `exec`s the extracted definitions in memory,
and stores only inputs and outputs.

What is executed
  * `multi_scale_deformable_attn_pytorch`
        third_party/mmcv/mmcv/ops/multi_scale_deform_attn.py:92-149
    (forward, and backward through torch autograd) on
      - the seeded vectors of the reference's own tests
        (third_party/mmcv/tests/test_ops/test_ms_deformable_attn.py:54-70, 138-182),
      - PAVE-Net-shaped cases (8 heads x 32 channels, 4 levels, P = 4 / 15 / 17),
      - edge cases: locations outside [0,1], exactly on pixel centres / borders,
        1x1 levels, NaN / Inf locations.
  * the forward maths of the module classes that feed the op
        MultiScaleDeformableAttention                         multi_scale_deform_attn.py:207-412
        MulFramesMultiScaleDeformableAttentionNumFrames3/5    multi_scale_deform_attn.py:1268-1982
        MultiScaleDeformablePoseAttention                     opera/models/utils/transformer.py:251-427
        MulFramesMultiScaleDeformablePoseAttentionNumFrames3/5  transformer.py:1543-1863, 2738-3114
    with mmcv's base classes / registries stubbed (mmcv itself is not importable
    in this image), the CUDA op replaced by the reference's
    own CPU function above, and NumFrames3's debug visualisation
    (transformer.py:1817-1830) disabled.  The classes branch on
    `torch.cuda.is_available()`; it is forced True so the branch every GPU run
    takes (the one that is not broken) is the one recorded.
"""
import ast
import math
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import recipe256  # noqa: E402  (seeded parameters / inputs of the production-geometry cases)
MSDA_PY = os.path.join(REF, 'third_party/mmcv/mmcv/ops/multi_scale_deform_attn.py')
OT_PY = os.path.join(REF, 'opera/models/utils/transformer.py')


def extract(path, names):
    """Return {name: source} for top-level defs/classes `names` of file `path`."""
    src = open(path, encoding='utf-8').read()
    tree = ast.parse(src)
    found = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            seg = ast.get_source_segment(src, node)
            # decorators are not part of the segment for classes; we do not want them anyway
            found[node.name] = seg
    missing = set(names) - set(found)
    if missing:
        raise SystemExit('not found in %s: %s' % (path, sorted(missing)))
    return found


# --- 1. the op -------------------------------------------------------------
ns_op = {'torch': torch, 'F': F}
exec(extract(MSDA_PY, ['multi_scale_deformable_attn_pytorch'])
     ['multi_scale_deformable_attn_pytorch'], ns_op)
ref_op = ns_op['multi_scale_deformable_attn_pytorch']


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def run_op_case(name, value, shapes, loc, aw, grad_seed, store):
    value = value.clone().requires_grad_(True)
    loc = loc.clone().requires_grad_(True)
    aw = aw.clone().requires_grad_(True)
    out = ref_op(value, shapes, loc, aw)
    g = torch.Generator().manual_seed(grad_seed)
    grad_out = torch.randn(out.shape, generator=g, dtype=out.dtype)
    out.backward(grad_out)
    store[name + '.value'] = value.detach().numpy()
    store[name + '.shapes'] = shapes.numpy()
    store[name + '.loc'] = loc.detach().numpy()
    store[name + '.aw'] = aw.detach().numpy()
    store[name + '.out'] = out.detach().numpy()
    store[name + '.grad_out'] = grad_out.numpy()
    store[name + '.grad_value'] = value.grad.numpy()
    # NaN / Inf locations: autograd of grid_sample yields NaN there; the CUDA
    # semantics (sample skipped) are pinned on the forward only for such cases
    store[name + '.grad_loc'] = loc.grad.numpy()
    store[name + '.grad_aw'] = aw.grad.numpy()


def mmcv_test_inputs(dtype):
    """test_ms_deformable_attn.py:54-66 (seed 3)."""
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = sum((H * W).item() for H, W in shapes)
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    sampling_locations = torch.rand(N, Lq, M, L, P, 2)
    attention_weights = torch.rand(N, Lq, M, L, P) + 1e-5
    attention_weights /= attention_weights.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value.to(dtype), shapes, sampling_locations.to(dtype), attention_weights.to(dtype)


def gradcheck_inputs(channels, seed):
    """test_ms_deformable_attn.py:150-162 (unseeded there; seeded here)."""
    N, M = 1, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(3, 2), (2, 1)], dtype=torch.long)
    S = sum((H * W).item() for H, W in shapes)
    torch.manual_seed(seed)
    value = torch.rand(N, S, M, channels) * 0.01
    sampling_locations = torch.rand(N, Lq, M, L, P, 2)
    attention_weights = torch.rand(N, Lq, M, L, P) + 1e-5
    attention_weights /= attention_weights.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value.double(), shapes, sampling_locations.double(), attention_weights.double()


def pavenet_inputs(seed, B, Q, P, shapes, dtype, M=8, D=32, spread=0.15):
    """PAVE-Net head geometry (8 heads x 32 channels, 4 levels); locations
    spill `spread` outside [0,1] on every side so the zero-padding and
    partial-corner paths are exercised."""
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    g = torch.Generator().manual_seed(seed)
    value = torch.randn(B, S, M, D, generator=g)
    loc = torch.rand(B, Q, M, L, P, 2, generator=g) * (1 + 2 * spread) - spread
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P)
    return value.to(dtype), shapes, loc.to(dtype), aw.to(dtype)


def edge_inputs(dtype):
    """Hand-placed locations: pixel centres, map borders, the (-1,0) band,
    exactly -1 / H (excluded), far outside, and a 1x1 level."""
    shapes = torch.as_tensor([(4, 5), (1, 1), (2, 3)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    M, D, L = 2, 4, 3
    g = torch.Generator().manual_seed(11)
    value = torch.randn(1, S, M, D, generator=g)
    pts = []
    for H, W in shapes.tolist():
        lvl = [
            (0.5 / W, 0.5 / H),            # centre of pixel (0,0): lh = lw = 0
            ((W - 0.5) / W, (H - 0.5) / H),  # centre of the last pixel
            (0.0, 0.0),                    # h = w = -0.5: only corner 4 valid
            (1.0, 1.0),                    # h = H-0.5: only corner 1 valid
            (-0.5 / W, 0.3),               # w = -1 exactly: excluded
            ((W + 0.5) / W, 0.3),          # w = W exactly: excluded
            (-0.4 / W, 0.5),               # w in (-1, -0.5)
            (0.5, (H + 0.4) / H),          # h in (H-0.5, H)
            (3.0, -2.0),                   # far outside
            (0.37, 0.81),                  # interior
        ]
        pts.append(lvl)
    P = len(pts[0])
    loc = torch.tensor(pts, dtype=torch.float64)          # (L, P, 2)
    loc = loc[None, None, None].repeat(1, 3, M, 1, 1, 1)   # (1, Q=3, M, L, P, 2)
    loc[:, 1] += 1e-3                                      # a second query, nudged
    loc[:, 2] -= 1e-3
    aw = torch.softmax(torch.randn(1, 3, M, L * P, generator=g), -1).view(1, 3, M, L, P)
    return value.to(dtype), shapes, loc.to(dtype), aw.to(dtype)


def gen_op_golden():
    store = {}
    run_op_case('mmcv_f64', *mmcv_test_inputs(torch.float64), 1, store)
    run_op_case('mmcv_f32', *mmcv_test_inputs(torch.float32), 1, store)
    for i, ch in enumerate([4, 30, 32, 64, 71, 1025]):
        run_op_case('gradcheck_c%d' % ch, *gradcheck_inputs(ch, 100 + i), 2, store)
    small = [(8, 12), (4, 6), (2, 3), (1, 2)]
    run_op_case('enc_f32', *pavenet_inputs(5, 2, 21, 4, small, torch.float32), 3, store)
    run_op_case('enc_f64', *pavenet_inputs(5, 1, 9, 4, small, torch.float64, M=2), 3, store)
    run_op_case('pose17_f32', *pavenet_inputs(6, 1, 5, 17, small, torch.float32), 4, store)
    run_op_case('pose15_f64', *pavenet_inputs(7, 1, 3, 15, small, torch.float64, M=2), 5, store)
    run_op_case('d16_f32', *pavenet_inputs(8, 1, 11, 4, small, torch.float32, M=4, D=16), 6, store)
    run_op_case('d64_f32', *pavenet_inputs(9, 1, 7, 3, small, torch.float32, M=2, D=64), 7, store)
    run_op_case('edge_f64', *edge_inputs(torch.float64), 8, store)
    run_op_case('edge_f32', *edge_inputs(torch.float32), 8, store)
    # NaN / Inf locations: forward only (sample contributes nothing on the GPU;
    # grid_sample propagates NaN, so only the finite queries are comparable)
    v, s, loc, aw = pavenet_inputs(10, 1, 4, 4, small, torch.float32, M=2)
    out_clean = ref_op(v, s, loc, aw)
    store['nonfinite_f32.value'] = v.numpy()
    store['nonfinite_f32.shapes'] = s.numpy()
    store['nonfinite_f32.loc'] = loc.numpy()
    store['nonfinite_f32.aw'] = aw.numpy()
    store['nonfinite_f32.out_clean'] = out_clean.numpy()
    np.savez_compressed(os.path.join(OUT, 'op_golden.npz'), **store)
    print('op_golden.npz: %d arrays' % len(store))


# --- 2. the module classes ---------------------------------------------------
class _TorchProxy(types.ModuleType):
    """`torch`, except that torch.cuda.is_available() says True (see module docstring)."""

    def __init__(self):
        super().__init__('torch')
        cuda = types.SimpleNamespace(is_available=lambda: True)
        self.__dict__['cuda'] = cuda

    def __getattr__(self, name):
        return getattr(torch, name)


class _CudaLike(torch.Tensor):
    """A CPU tensor that answers `is_cuda == True`, so that the classes whose
    dispatch reads `torch.cuda.is_available() and value.is_cuda`
    (multi_scale_deform_attn.py:398, 1550, 1925) take their GPU branch."""

    @property
    def is_cuda(self):
        return True


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def _constant_init(module, val, bias=0):
    nn.init.constant_(module.weight, val)
    if module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _xavier_init(module, gain=1, bias=0, distribution='normal'):
    (nn.init.xavier_uniform_ if distribution == 'uniform' else nn.init.xavier_normal_)(
        module.weight, gain=gain)
    if module.bias is not None:
        nn.init.constant_(module.bias, bias)


class _FunctionStub:
    @staticmethod
    def apply(value, shapes, lsi, loc, aw, im2col_step):
        return ref_op(value, shapes, loc, aw)


def _op_any_arity(*args):
    if len(args) == 6:   # the reference's broken CPU-branch call shape
        return ref_op(args[0], args[1], args[3], args[4])
    return ref_op(*args)


def module_namespace():
    return {
        'torch': _TorchProxy(), 'nn': nn, 'F': F, 'math': math, 'warnings': warnings,
        'BaseModule': _BaseModule, 'ATTENTION': _Registry(),
        'deprecated_api_warning': lambda *a, **k: (lambda fn: fn),
        'constant_init': _constant_init, 'xavier_init': _xavier_init,
        'MultiScaleDeformableAttnFunction': _FunctionStub,
        'multi_scale_deformable_attn_pytorch': _op_any_arity,
        'print': lambda *a, **k: None,
    }


def randomise(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            scale = 0.5 if 'sampling_offsets' in name else 0.3
            p.copy_(torch.randn(p.shape, generator=g) * scale)


def save_module_case(store, name, module, cfg, inputs, out):
    for k, v in module.state_dict().items():
        store['%s.state.%s' % (name, k)] = v.numpy()
    for k, v in cfg.items():
        store['%s.cfg.%s' % (name, k)] = np.asarray(v)
    for k, v in inputs.items():
        if v is not None:
            store['%s.in.%s' % (name, k)] = v.numpy()
    store['%s.out' % name] = out.detach().numpy()


def gen_module_golden():
    store = {}
    ns_m = module_namespace()
    for cls_name, src in extract(MSDA_PY, [
            'MultiScaleDeformableAttention',
            'MulFramesMultiScaleDeformableAttentionNumFrames3',
            'MulFramesMultiScaleDeformableAttentionNumFrames5']).items():
        exec(src, ns_m)
    ns_o = module_namespace()
    for cls_name, src in extract(OT_PY, [
            'MultiScaleDeformablePoseAttention',
            'MulFramesMultiScaleDeformablePoseAttentionNumFrames3',
            'MulFramesMultiScaleDeformablePoseAttentionNumFrames5']).items():
        exec(src, ns_o)

    C, M, L = 32, 4, 3
    shapes = torch.as_tensor([(6, 9), (3, 5), (2, 2)], dtype=torch.long)
    lsi = lsi_of(shapes)
    S = int(shapes.prod(1).sum())
    g = torch.Generator().manual_seed(2024)

    def rnd(*shape):
        return torch.randn(*shape, generator=g)

    def mask(*shape):
        m = torch.rand(*shape, generator=g) < 0.15
        return m

    # ---- encoder self-attention (MultiScaleDeformableAttention) ----
    cfg = dict(embed_dims=C, num_heads=M, num_levels=L, num_points=4)
    mod = ns_m['MultiScaleDeformableAttention'](dropout=0.0, **cfg).eval()
    randomise(mod, 1)
    B = 2
    inp = dict(query=rnd(S, B, C), query_pos=rnd(S, B, C), key_padding_mask=mask(B, S),
               reference_points=torch.rand(B, S, L, 2, generator=g))
    out = mod(inp['query'], None, None, query_pos=inp['query_pos'],
              key_padding_mask=inp['key_padding_mask'], reference_points=inp['reference_points'],
              spatial_shapes=shapes, level_start_index=lsi)
    inp['spatial_shapes'] = shapes
    save_module_case(store, 'encoder', mod, cfg, inp, out)
    # 4-d reference boxes branch (multi_scale_deform_attn.py:389-393)
    inp4 = dict(query=rnd(7, B, C), value=rnd(S, B, C),
                reference_points=torch.rand(B, 7, L, 4, generator=g))
    out = mod(inp4['query'], None, inp4['value'], reference_points=inp4['reference_points'],
              spatial_shapes=shapes, level_start_index=lsi)
    inp4['spatial_shapes'] = shapes
    save_module_case(store, 'encoder_box', mod, cfg, inp4, out)

    # ---- PETR pose attention ----
    K = 5
    cfg = dict(embed_dims=C, num_heads=M, num_levels=L, num_points=K)
    mod = ns_o['MultiScaleDeformablePoseAttention'](dropout=0.0, **cfg).eval()
    randomise(mod, 2)
    Qp = 6
    inp = dict(query=rnd(Qp, B, C), query_pos=rnd(Qp, B, C), value=rnd(S, B, C),
               key_padding_mask=mask(B, S),
               reference_points=torch.rand(B, Qp, L, 2 * K, generator=g))
    out = mod(inp['query'], None, inp['value'], query_pos=inp['query_pos'],
              key_padding_mask=inp['key_padding_mask'], reference_points=inp['reference_points'],
              spatial_shapes=shapes, level_start_index=lsi)
    inp['spatial_shapes'] = shapes
    save_module_case(store, 'pose', mod, cfg, inp, out)

    # ---- multi-frame pose-aware attention, T = 3 and 5 ----
    for T, cls in ((3, 'MulFramesMultiScaleDeformablePoseAttentionNumFrames3'),
                   (5, 'MulFramesMultiScaleDeformablePoseAttentionNumFrames5')):
        cfg = dict(embed_dims=C, num_heads=M, num_levels=L, num_points=K)
        mod = ns_o[cls](dropout=0.0, **cfg).eval()
        if hasattr(mod, 'vis_attention'):
            mod.vis_attention = lambda *a, **k: None   # debug code, transformer.py:1817-1830
        randomise(mod, 10 + T)
        Bc = 2
        inp = dict(query=rnd(Qp, Bc, C), query_pos=rnd(Qp, Bc, C), value=rnd(S, Bc * T, C),
                   key_padding_mask=mask(Bc * T, S),
                   reference_points=torch.rand(Bc, T * Qp, L, 2 * K, generator=g))
        out = mod(inp['query'], None, inp['value'], query_pos=inp['query_pos'],
                  key_padding_mask=inp['key_padding_mask'],
                  reference_points=inp['reference_points'], spatial_shapes=shapes,
                  level_start_index=lsi)
        inp['spatial_shapes'] = shapes
        cfg['num_frames'] = T
        save_module_case(store, 'mf_pose%d' % T, mod, cfg, inp, out)

    # ---- multi-frame joint-decoder attention, T = 3 and 5 ----
    for T, cls in ((3, 'MulFramesMultiScaleDeformableAttentionNumFrames3'),
                   (5, 'MulFramesMultiScaleDeformableAttentionNumFrames5')):
        cfg = dict(embed_dims=C, num_heads=M, num_levels=L, num_points=4)
        mod = ns_m[cls](dropout=0.0, **cfg).eval()
        randomise(mod, 20 + T)
        G, Qj = 3, K
        inp = dict(query=rnd(Qj, G, C), query_pos=rnd(Qj, G, C), value=rnd(S, G, T, C),
                   key_padding_mask=mask(G, T, S),
                   reference_points=torch.rand(T * G, Qj, L, 2, generator=g))
        out = mod(inp['query'], None, inp['value'].as_subclass(_CudaLike),
                  query_pos=inp['query_pos'], key_padding_mask=inp['key_padding_mask'],
                  reference_points=inp['reference_points'], spatial_shapes=shapes,
                  level_start_index=lsi).as_subclass(torch.Tensor)
        inp['spatial_shapes'] = shapes
        cfg['num_frames'] = T
        save_module_case(store, 'mf_joint%d' % T, mod, cfg, inp, out)

    np.savez_compressed(os.path.join(OUT, 'module_golden.npz'), **store)
    print('module_golden.npz: %d arrays' % len(store))


def gen_module_golden_256():
    """The six module classes at PAVE-Net's PRODUCTION geometry — embed_dims 256, 8 heads
    (32 channels per head), 4 levels, P = 4 / 15 / 17
    (configs/videopose/2025-2-13/2025_2_13_res50_num_frames_3_posetrack17.py:53-107) — forward
    AND backward (torch autograd through the reference classes).  Parameters and inputs are
    NOT stored: tests/golden/recipe256.py regenerates them from seeds (the state dicts alone
    would be ~20 MB); what is stored is the reference's output, the gradients of the inputs,
    the bias gradients, and two seeded random projections of every weight gradient
    (u^T dW and dW v), plus a checksum of the regenerated tensors."""
    store = {}
    ns_m = module_namespace()
    for _, src in extract(MSDA_PY, [
            'MultiScaleDeformableAttention',
            'MulFramesMultiScaleDeformableAttentionNumFrames3',
            'MulFramesMultiScaleDeformableAttentionNumFrames5']).items():
        exec(src, ns_m)
    ns_o = module_namespace()
    for _, src in extract(OT_PY, [
            'MultiScaleDeformablePoseAttention',
            'MulFramesMultiScaleDeformablePoseAttentionNumFrames3',
            'MulFramesMultiScaleDeformablePoseAttentionNumFrames5']).items():
        exec(src, ns_o)
    namespaces = {'msda': ns_m, 'ot': ns_o}
    for case in recipe256.CASES:
        cls = namespaces[case['where']][case['cls']]
        mod = cls(dropout=0.0, **case['cfg']).eval()
        if hasattr(mod, 'vis_attention'):
            mod.vis_attention = lambda *a, **k: None   # debug code, transformer.py:1817-1830
        recipe256.randomise(mod, case['param_seed'])
        inp = recipe256.make_inputs(case)
        leaves = {k: v.clone().requires_grad_(True) for k, v in inp.items()
                  if k in ('query', 'value', 'query_pos')}
        value = leaves.get('value')
        if value is not None and case.get('cuda_like'):
            value = value.as_subclass(_CudaLike)
        out = mod(leaves['query'], None, value, query_pos=leaves.get('query_pos'),
                  key_padding_mask=inp.get('key_padding_mask'),
                  reference_points=inp['reference_points'], spatial_shapes=inp['spatial_shapes'],
                  level_start_index=lsi_of(inp['spatial_shapes']))
        out = out.as_subclass(torch.Tensor)
        grad_out = recipe256.grad_output(case, out.shape)
        out.backward(grad_out)
        name = case['name']
        store[name + '.out'] = out.detach().numpy()
        for k, v in leaves.items():
            store['%s.grad_in.%s' % (name, k)] = v.grad.as_subclass(torch.Tensor).numpy()
        for pname, p in sorted(mod.named_parameters()):
            g = p.grad
            if g.dim() == 1:
                store['%s.grad_param.%s' % (name, pname)] = g.numpy()
            else:
                u, v = recipe256.projection_vectors(case, pname, g.shape)
                store['%s.grad_param_u.%s' % (name, pname)] = (u @ g).numpy()
                store['%s.grad_param_v.%s' % (name, pname)] = (g @ v).numpy()
        store[name + '.checksum'] = recipe256.checksum(mod, inp).numpy()
    np.savez_compressed(os.path.join(OUT, 'module_golden_256.npz'), **store)
    print('module_golden_256.npz: %d arrays' % len(store))


if __name__ == '__main__':
    torch.set_num_threads(4)
    gen_op_golden()
    gen_module_golden()
    gen_module_golden_256()
