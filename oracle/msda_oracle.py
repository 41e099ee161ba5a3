"""CPU ORACLE for the multi-scale deformable attention hot path.

TEST INFRASTRUCTURE — not product code.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module; `pavenet_b200/` never does.

Parity status: PINNED.  `tests/test_oracle.py` checks everything here against
golden vectors generated in the build container by executing the reference's
own code (`tests/golden/gen_golden.py`: the function
`multi_scale_deformable_attn_pytorch`, third_party/mmcv/mmcv/ops/multi_scale_deform_attn.py:92-149,
and the module classes' forward maths, extracted from the reference checkout),
including the seeded vectors of the reference's own test-suite
(third_party/mmcv/tests/test_ops/test_ms_deformable_attn.py:54-182).

Contents
  * `c_forward` / `c_backward`      ctypes front-end of oracle/msda_ref.c (plain-C loops
                                     following ms_deform_attn_cuda_kernel.cuh:17-131,200-254)
  * `grid_sample_port`              restatement of the reference's CPU path
                                     (per-level F.grid_sample, multi_scale_deform_attn.py:92-149);
                                     differentiable through autograd — the CPU baseline the
                                     bench times
  * `*_attention_ref`               restatements of the module-level maths around the op
                                     (multi_scale_deform_attn.py:353-412, 1437-1587;
                                     opera/models/utils/transformer.py:368-427, 1685-1863, 2893-3114)
"""
import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libmsda_ref.so')
_lib = None


def build(force=False):
    """Compile oracle/msda_ref.c -> oracle/libmsda_ref.so (gcc via the Makefile)."""
    src = os.path.join(_HERE, 'msda_ref.c')
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    proc = subprocess.run(['make', '-C', _HERE, '-B', 'libmsda_ref.so'],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('building the C oracle failed:\n' + proc.stdout)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for name in ('msda_ref_forward_f32', 'msda_ref_forward_f64'):
            getattr(_lib, name).restype = None
            getattr(_lib, name).argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 7
        for name in ('msda_ref_backward_f32', 'msda_ref_backward_f64'):
            getattr(_lib, name).restype = None
            getattr(_lib, name).argtypes = [ctypes.c_void_p] * 9 + [ctypes.c_int] * 7
    return _lib


def _prep(value, shapes, lsi, loc, aw):
    dt = loc.dtype
    assert dt in (torch.float32, torch.float64)
    value = value.detach().to('cpu', dt).contiguous()
    loc = loc.detach().to('cpu', dt).contiguous()
    aw = aw.detach().to('cpu', dt).contiguous()
    shapes = shapes.detach().to('cpu', torch.int64).contiguous()
    if lsi is None:
        lsi = level_start_index(shapes)
    lsi = lsi.detach().to('cpu', torch.int64).contiguous()
    B, S, M, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    return value, shapes, lsi, loc, aw, (B, S, M, D, L, Q, P)


def level_start_index(shapes):
    """[0, H0*W0, H0*W0+H1*W1, ...] as the callers build it
    (e.g. opera/models/utils/transformer.py `level_start_index` construction)."""
    sizes = shapes[:, 0] * shapes[:, 1]
    return torch.cat([sizes.new_zeros(1), sizes.cumsum(0)[:-1]])


def c_forward(value, shapes, lsi, loc, aw):
    """Forward through the plain-C loops.  Returns (B, Q, M*D) on the CPU."""
    value, shapes, lsi, loc, aw, dims = _prep(value, shapes, lsi, loc, aw)
    B, S, M, D, L, Q, P = dims
    out = torch.empty((B, Q, M * D), dtype=loc.dtype)
    fn = _load().msda_ref_forward_f32 if loc.dtype == torch.float32 else _load().msda_ref_forward_f64
    fn(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), aw.data_ptr(),
       out.data_ptr(), B, S, M, D, L, Q, P)
    return out


def c_backward(value, shapes, lsi, loc, aw, grad_out):
    """Backward through the plain-C loops.  Returns (grad_value, grad_loc, grad_aw)."""
    value, shapes, lsi, loc, aw, dims = _prep(value, shapes, lsi, loc, aw)
    B, S, M, D, L, Q, P = dims
    grad_out = grad_out.detach().to('cpu', loc.dtype).contiguous()
    gv = torch.zeros_like(value)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(aw)
    fn = _load().msda_ref_backward_f32 if loc.dtype == torch.float32 else _load().msda_ref_backward_f64
    fn(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), aw.data_ptr(),
       grad_out.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), B, S, M, D, L, Q, P)
    return gv, gl, ga


def grid_sample_port(value, shapes, loc, aw):
    """The reference's CPU path, restated: one `F.grid_sample` (bilinear, zero
    padding, align_corners=False) per level on the grid 2*loc-1, then the
    attention-weighted sum over levels and points.

    value (B,S,M,D), shapes (L,2) [H,W], loc (B,Q,M,L,P,2) in (x,y), aw (B,Q,M,L,P)
    -> (B, Q, M*D).  Differentiable (autograd supplies the backward the bench
    times as the CPU baseline).
    """
    B, S, M, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = [(int(h), int(w)) for h, w in shapes.tolist()]
    # heads become grid_sample's batch: (B*M, D, S)
    per_head = value.permute(0, 2, 3, 1).reshape(B * M, D, S)
    grids = (loc * 2 - 1).permute(0, 2, 1, 3, 4, 5).reshape(B * M, Q, L, P, 2)
    sampled = []
    start = 0
    for lvl, (h, w) in enumerate(hw):
        fmap = per_head[:, :, start:start + h * w].reshape(B * M, D, h, w)
        start += h * w
        sampled.append(F.grid_sample(fmap, grids[:, :, lvl], mode='bilinear',
                                     padding_mode='zeros', align_corners=False))
    sampled = torch.stack(sampled, dim=3)                       # (B*M, D, Q, L, P)
    weights = aw.permute(0, 2, 1, 3, 4).reshape(B * M, 1, Q, L, P)
    out = (sampled * weights).sum(dim=(3, 4))                   # (B*M, D, Q)
    return out.reshape(B, M * D, Q).transpose(1, 2).contiguous()


# ---------------------------------------------------------------------------
# module-level compositions (what the callers compute around the op)
# ---------------------------------------------------------------------------
def _lin(state, name, x):
    return F.linear(x, state[name + '.weight'], state[name + '.bias'])


def _pose_box_wh(ref_kpts):
    """(..., L, 2K) keypoint reference -> (..., L, 2) clamped pose-box (w, h)
    (opera/models/utils/transformer.py:402-410)."""
    xs, ys = ref_kpts[..., 0::2], ref_kpts[..., 1::2]
    w = (xs.max(-1, keepdim=True)[0] - xs.min(-1, keepdim=True)[0]).clamp(min=1e-4)
    h = (ys.max(-1, keepdim=True)[0] - ys.min(-1, keepdim=True)[0]).clamp(min=1e-4)
    return torch.cat([w, h], -1)


def encoder_attention_ref(state, cfg, query, value=None, identity=None, query_pos=None,
                          key_padding_mask=None, reference_points=None, spatial_shapes=None,
                          op=grid_sample_port):
    """`MultiScaleDeformableAttention.forward` in eval mode (no dropout),
    seq-first I/O (multi_scale_deform_attn.py:353-412)."""
    M, L, P = cfg['num_heads'], cfg['num_levels'], cfg['num_points']
    value = query if value is None else value
    identity = query if identity is None else identity
    if query_pos is not None:
        query = query + query_pos
    query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
    B, Q, C = query.shape
    S = value.shape[1]
    v = _lin(state, 'value_proj', value)
    if key_padding_mask is not None:
        v = v.masked_fill(key_padding_mask[..., None], 0.0)
    v = v.view(B, S, M, C // M)
    off = _lin(state, 'sampling_offsets', query).view(B, Q, M, L, P, 2)
    w = _lin(state, 'attention_weights', query).view(B, Q, M, L * P).softmax(-1).view(B, Q, M, L, P)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([spatial_shapes[:, 1], spatial_shapes[:, 0]], -1).to(off.dtype)
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = (reference_points[:, :, None, :, None, :2]
               + off / P * reference_points[:, :, None, :, None, 2:] * 0.5)
    out = op(v, spatial_shapes, loc, w)
    return _lin(state, 'output_proj', out).permute(1, 0, 2) + identity


def pose_attention_ref(state, cfg, query, value, query_pos=None, key_padding_mask=None,
                       reference_points=None, spatial_shapes=None, op=grid_sample_port):
    """`MultiScaleDeformablePoseAttention.forward` in eval mode
    (opera/models/utils/transformer.py:368-427); reference_points (B,Q,L,2K)."""
    M, L, P = cfg['num_heads'], cfg['num_levels'], cfg['num_points']
    residual = query
    if query_pos is not None:
        query = query + query_pos
    query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
    B, Q, C = query.shape
    S = value.shape[1]
    v = _lin(state, 'value_proj', value)
    if key_padding_mask is not None:
        v = v.masked_fill(key_padding_mask[..., None], 0.0)
    v = v.view(B, S, M, C // M)
    off = _lin(state, 'sampling_offsets', query).view(B, Q, M, L, P, 2)
    w = _lin(state, 'attention_weights', query).view(B, Q, M, L * P).softmax(-1).view(B, Q, M, L, P)
    kp = reference_points.reshape(B, Q, L, P, 2)[:, :, None]
    wh = _pose_box_wh(reference_points)[:, :, None, :, None, :]
    loc = kp + off * wh * 0.5
    out = op(v, spatial_shapes, loc, w)
    return _lin(state, 'output_proj', out).permute(1, 0, 2) + residual


def frame_prefixes(num_frames):
    """Parameter-name prefixes per frame, oldest first
    (opera/models/utils/transformer.py:1611-1624 and 2802-2836)."""
    return {3: ['pre_', '', 'next_'],
            5: ['pre_pre_', 'pre_', '', 'next_', 'next_next_']}[num_frames]


def mulframes_pose_attention_ref(state, cfg, query, value, query_pos=None, key_padding_mask=None,
                                 reference_points=None, spatial_shapes=None,
                                 op=grid_sample_port):
    """`MulFramesMultiScaleDeformablePoseAttentionNumFrames{3,5}.forward`, eval
    mode, visualisation side-effects dropped (transformer.py:1685-1863, 2893-3114):
    mask BEFORE value_proj, frame t = value[t::T], T op calls, outputs fused
    with Z_t / sum Z where Z_t = sum exp(logits_t) (no max-subtraction).

    query (Q, Bc, C); value (S, Bc*T, C); reference_points (Bc, T*Q, L, 2K).
    """
    M, L, P, T = cfg['num_heads'], cfg['num_levels'], cfg['num_points'], cfg['num_frames']
    residual = query
    if query_pos is not None:
        query = query + query_pos
    query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
    Bc, Q, C = query.shape
    S = value.shape[1]
    if key_padding_mask is not None:
        value = value.masked_fill(key_padding_mask[..., None], 0.0)
    v_all = _lin(state, 'value_proj', value)
    outs, zs = [], []
    for t, pre in enumerate(frame_prefixes(T)):
        v = v_all[t::T].reshape(Bc, S, M, C // M)
        off = _lin(state, pre + 'sampling_offsets', query).view(Bc, Q, M, L, P, 2)
        logits = _lin(state, pre + 'attention_weights', query).view(Bc, Q, M, L * P)
        zs.append(torch.exp(logits).sum(-1, keepdim=True))
        w = logits.softmax(-1).view(Bc, Q, M, L, P)
        ref_t = reference_points[:, t * Q:(t + 1) * Q]
        kp = ref_t.reshape(Bc, Q, L, P, 2)[:, :, None]
        wh = _pose_box_wh(ref_t)[:, :, None, :, None, :]
        loc = kp + off * wh * 0.5
        outs.append(op(v, spatial_shapes, loc, w).reshape(Bc, Q, M, C // M))
    z_all = sum(zs)
    fused = sum(o * (z / z_all) for o, z in zip(outs, zs)).flatten(-2)
    return _lin(state, 'output_proj', fused).permute(1, 0, 2) + residual


def mulframes_joint_attention_ref(state, cfg, query, value, identity=None, query_pos=None,
                                  key_padding_mask=None, reference_points=None,
                                  spatial_shapes=None, op=grid_sample_port):
    """`MulFramesMultiScaleDeformableAttentionNumFrames{3,5}.forward`, eval mode
    (multi_scale_deform_attn.py:1437-1587, 1780-1982), 2-d reference points.

    query (Q, G, C); value (S, G, T, C); key_padding_mask (G, T, S);
    reference_points (T*G, Q, L, 2).
    """
    M, L, P, T = cfg['num_heads'], cfg['num_levels'], cfg['num_points'], cfg['num_frames']
    identity = query if identity is None else identity
    if query_pos is not None:
        query = query + query_pos
    query, value = query.permute(1, 0, 2), value.permute(1, 0, 2, 3)
    G, Q, C = query.shape
    S = value.shape[1]
    if key_padding_mask is not None:
        value = value.masked_fill(key_padding_mask.transpose(1, 2)[..., None], 0.0)
    v_all = _lin(state, 'value_proj', value)
    norm = torch.stack([spatial_shapes[:, 1], spatial_shapes[:, 0]], -1).to(query.dtype)
    outs, zs = [], []
    for t, pre in enumerate(frame_prefixes(T)):
        v = v_all[:, :, t].reshape(G, S, M, C // M)
        off = _lin(state, pre + 'sampling_offsets', query).view(G, Q, M, L, P, 2)
        logits = _lin(state, pre + 'attention_weights', query).view(G, Q, M, L * P)
        zs.append(torch.exp(logits).sum(-1, keepdim=True))
        w = logits.softmax(-1).view(G, Q, M, L, P)
        ref_t = reference_points[t * G:(t + 1) * G]
        loc = ref_t[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        outs.append(op(v, spatial_shapes, loc, w).reshape(G, Q, M, C // M))
    z_all = sum(zs)
    fused = sum(o * (z / z_all) for o, z in zip(outs, zs)).flatten(-2)
    return _lin(state, 'output_proj', fused).permute(1, 0, 2) + identity


# ---------------------------------------------------------------------------
# the REFERENCE's own CUDA kernels (oracle/_ref/libmsda_refcuda.so)
# ---------------------------------------------------------------------------
_REFCUDA_PATH = os.path.join(_HERE, '_ref', 'libmsda_refcuda.so')
_refcuda = None


def build_refcuda(reference='/root/reference', force=False):
    """Compile the reference's `ms_deform_attn_cuda_kernel.cuh` (from the reference checkout,
    nothing copied) behind `refcuda_driver.cu` into oracle/_ref/.  Build container only; on
    the GPU box the prebuilt library travels with the snapshot.  Returns the path or None when
    there is no reference checkout."""
    hdr = os.path.join(reference, 'third_party/mmcv/mmcv/ops/csrc/common/cuda/'
                                  'ms_deform_attn_cuda_kernel.cuh')
    if not os.path.exists(hdr):
        return _REFCUDA_PATH if os.path.exists(_REFCUDA_PATH) else None
    drv = os.path.join(_HERE, 'refcuda_driver.cu')
    if (not force and os.path.exists(_REFCUDA_PATH)
            and os.path.getmtime(_REFCUDA_PATH) >= os.path.getmtime(drv)):
        return _REFCUDA_PATH
    proc = subprocess.run(['make', '-C', _HERE, '-B', 'refcuda', 'REF=' + reference],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError('building the reference CUDA kernels failed:\n' + proc.stdout)
    return _REFCUDA_PATH


def refcuda_available():
    return os.path.exists(_REFCUDA_PATH)


def _load_refcuda():
    global _refcuda
    if _refcuda is None:
        lib = ctypes.CDLL(_REFCUDA_PATH)
        lib.refcuda_forward.restype = ctypes.c_int
        lib.refcuda_forward.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 8 + [ctypes.c_void_p]
        lib.refcuda_backward.restype = ctypes.c_int
        lib.refcuda_backward.argtypes = [ctypes.c_void_p] * 9 + [ctypes.c_int] * 8 + [ctypes.c_void_p]
        _refcuda = lib
    return _refcuda


def _refcuda_dims(value, loc):
    B, S, M, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    code = {torch.float32: 0, torch.float64: 1}[value.dtype]
    return (B, S, M, D, L, Q, P, code)


def refcuda_forward(value, shapes, lsi, loc, aw, out=None):
    """The reference's forward kernel (ms_deformable_im2col_gpu_kernel) on CUDA tensors,
    launched as the reference's host code does; enqueued on torch's current stream."""
    lib = _load_refcuda()
    B, S, M, D, L, Q, P, code = _refcuda_dims(value, loc)
    if out is None:
        out = torch.empty((B, Q, M * D), dtype=value.dtype, device=value.device)
    rc = lib.refcuda_forward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                             aw.data_ptr(), out.data_ptr(), B, S, M, D, L, Q, P, code,
                             torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError('refcuda_forward: CUDA error %d' % -rc)
    return out


def refcuda_backward(value, shapes, lsi, loc, aw, grad_out, grad_value, grad_loc, grad_aw):
    """The reference's backward kernel (ms_deformable_col2im_gpu_kernel_*, picked by channel count
    as ms_deform_attn_cuda.cu:63-206 does); ACCUMULATES into all three caller-zeroed gradients."""
    lib = _load_refcuda()
    B, S, M, D, L, Q, P, code = _refcuda_dims(value, loc)
    rc = lib.refcuda_backward(value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                              aw.data_ptr(), grad_out.data_ptr(), grad_value.data_ptr(),
                              grad_loc.data_ptr(), grad_aw.data_ptr(), B, S, M, D, L, Q, P, code,
                              torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError('refcuda_backward: CUDA error %d' % -rc)
