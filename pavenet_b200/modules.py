"""Host-side mirror of the reference's attention modules around the op.

Same class names, constructor kwargs, parameter (state-dict) names, init and
`forward` signatures as the reference, so configs under `configs/videopose`
and `configs/petr` build them unchanged:

  mmcv scope   MultiScaleDeformableAttention                       multi_scale_deform_attn.py:207-412
               MulFramesMultiScaleDeformableAttentionNumFrames3    multi_scale_deform_attn.py:1268-1587
               MulFramesMultiScaleDeformableAttentionNumFrames5    multi_scale_deform_attn.py:1590-1982
  opera scope  MultiScaleDeformablePoseAttention                   opera/models/utils/transformer.py:251-427
               MulFramesMultiScaleDeformablePoseAttentionNumFrames3  transformer.py:1543-1863
               MulFramesMultiScaleDeformablePoseAttentionNumFrames5  transformer.py:2738-3114

What is different underneath:
  * the op is the sm_100a kernel behind the C ABI (no CPU branch: a CPU tensor
    raises, it does not silently run PyTorch);
  * the multi-frame classes issue ONE op call per forward instead of T: the T
    frames of a clip become T*L "levels" of one value tensor and the per-frame
    softmax + Z_t/sum(Z) fusion of the reference (transformer.py:1736-1745,
    1854-1858) becomes one joint softmax over T*L*P, which is the same function
    wherever the reference is finite and does not
    overflow where the reference does.  `fused=False` reproduces the reference's
    T-call formulation literally (used by the parity tests);
  * the T offset / weight projections run as one GEMM each (weights
    concatenated on the fly; parameters stay separate, names unchanged);
  * NumFrames3's debug visualisation (transformer.py:1817-1830) is not carried over.
"""
import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from .functional import (FusedMultiScaleDeformableAttnFunction, MultiScaleDeformableAttnFunction,
                         ffn_supported, fuse_frames_as_levels, fused_ffn, fused_supported, layer_norm256,
                         layer_norm_supported, linear256, linear256_supported)
from .registry import ATTENTION, FEEDFORWARD_NETWORK, OPERA_ATTENTION

__all__ = [
    'MultiScaleDeformableAttention', 'MultiScaleDeformablePoseAttention',
    'MulFramesMultiScaleDeformablePoseAttentionNumFrames3',
    'MulFramesMultiScaleDeformablePoseAttentionNumFrames5',
    'MulFramesMultiScaleDeformableAttentionNumFrames3',
    'MulFramesMultiScaleDeformableAttentionNumFrames5', 'FFN', 'LayerNorm',
]

_FRAME_PREFIXES = {3: ('pre_', '', 'next_'),
                   5: ('pre_pre_', 'pre_', '', 'next_', 'next_next_')}


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError('invalid input for _is_power_of_2: {} (type: {})'.format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


def _ring_offsets(num_heads, num_levels, num_points):
    """Initial sampling-offset bias: head m looks along direction 2*pi*m/M,
    point p at p+1 pixels (multi_scale_deform_attn.py:286-297)."""
    thetas = torch.arange(num_heads, dtype=torch.float32) * (2.0 * math.pi / num_heads)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(num_heads, 1, 1, 2)
    grid = grid.repeat(1, num_levels, num_points, 1)
    for i in range(num_points):
        grid[:, :, i, :] *= i + 1
    return grid.view(-1)


def _const_linear(layer, weight=0., bias=0.):
    nn.init.constant_(layer.weight, weight)
    nn.init.constant_(layer.bias, bias)


def _xavier_linear(layer):
    nn.init.xavier_uniform_(layer.weight)
    nn.init.constant_(layer.bias, 0.)


def _run_op(value, spatial_shapes, level_start_index, loc, weights, im2col_step):
    if not value.is_cuda:
        raise RuntimeError('pavenet_b200 attention modules run on CUDA tensors only '
                           '(there is no CPU fallback); got value on %s' % value.device)
    return MultiScaleDeformableAttnFunction.apply(
        value.contiguous(), spatial_shapes, level_start_index, loc.contiguous(),
        weights.contiguous(), im2col_step)


def _run_fused(value, spatial_shapes, level_start_index, offsets, logits, ref_points, scale):
    """softmax + location transform + sampling in one kernel (forward and backward)."""
    return FusedMultiScaleDeformableAttnFunction.apply(
        value, spatial_shapes, level_start_index, offsets, logits, ref_points, scale)


class _DeformAttnBase(nn.Module):
    """Shared constructor logic (argument checks are the reference's)."""

    #: fold softmax and the reference-point transform into the sampling kernels
    #: when the shapes allow it (CUDA, fp32 projections, 32 channels per head);
    #: False reproduces the reference's op-by-op composition
    fuse_prologue = True

    #: run the 128/256-wide projections (value_proj, output_proj, and sampling_offsets /
    #: attention_weights when they have 128 or 256 outputs) on the tcgen05 tensor cores (3xTF32, fp32-level accuracy) with mask and
    #: storage dtype folded into the epilogue; False uses nn.Linear / cuBLAS fp32 op by op
    tensor_core_linear = True

    def _can_fuse(self, value, offsets):
        return self.fuse_prologue and fused_supported(value, offsets)

    def _project(self, layer, x, row_mask=None, mask_mode=0, out_dtype=None):
        """layer(x), with the padding mask applied after (mask_mode 1) or before
        (mask_mode 2) the projection and an optional storage dtype."""
        if self.tensor_core_linear and linear256_supported(x, layer.weight):
            return linear256(x, layer.weight, layer.bias, row_mask, mask_mode,
                             out_dtype or torch.float32)
        if row_mask is not None and mask_mode == 2:
            x = x.masked_fill(row_mask[..., None], 0.0)
        y = layer(x)
        if row_mask is not None and mask_mode == 1:
            y = y.masked_fill(row_mask[..., None], 0.0)
        if out_dtype is not None and y.dtype != out_dtype:
            y = y.to(out_dtype)
        return y

    def _project_out(self, output, identity, seq_first):
        """identity + dropout(output_proj(output)) for a batch-first `output`
        (multi_scale_deform_attn.py:406-412); seq_first: identity and the result are
        (Q, B, C).  Batch-first callers get dropout and the residual from the GEMM epilogue."""
        layer = self.output_proj
        if (not seq_first and self.tensor_core_linear and linear256_supported(output, layer.weight)
                and identity.dtype == torch.float32 and identity.is_cuda
                and identity.shape == output.shape[:-1] + (layer.out_features,)):
            return linear256(output, layer.weight, layer.bias, residual=identity,
                             dropout_p=self.dropout.p if self.training else 0.0)
        output = self._project(layer, output)
        if seq_first:
            output = output.permute(1, 0, 2)
        return self.dropout(output) + identity

    def __init__(self, embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                 batch_first, norm_cfg, init_cfg, value_dtype):
        super().__init__()
        if embed_dims % num_heads != 0:
            raise ValueError(f'embed_dims must be divisible by num_heads, '
                             f'but got {embed_dims} and {num_heads}')
        dim_per_head = embed_dims // num_heads
        if not _is_power_of_2(dim_per_head):
            warnings.warn("You'd better set embed_dims in MultiScaleDeformAttention to make "
                          'the dimension of each attention head a power of 2 '
                          'which is more efficient in our CUDA implementation.')
        self.norm_cfg = norm_cfg
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.batch_first = batch_first
        self.im2col_step = im2col_step
        self.embed_dims = embed_dims
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        #: None keeps value in the projection's dtype; torch.bfloat16 stores the
        #: projected value in bf16 (locations / weights / output stay fp32)
        self.value_dtype = value_dtype

    def _store(self, value):
        if self.value_dtype is not None and value.dtype != self.value_dtype:
            value = value.to(self.value_dtype)
        return value


@ATTENTION.register_module()
class MultiScaleDeformableAttention(_DeformAttnBase):
    """Deformable-DETR attention; PAVE-Net's spatial encoder self-attention.
    Mirrors multi_scale_deform_attn.py:207-412."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None, value_dtype=None):
        super().__init__(embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                         batch_first, norm_cfg, init_cfg, value_dtype)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        _const_linear(self.sampling_offsets, 0.)
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(
                _ring_offsets(self.num_heads, self.num_levels, self.num_points))
        _const_linear(self.attention_weights, 0., 0.)
        _xavier_linear(self.value_proj)
        _xavier_linear(self.output_proj)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        """query (num_query, bs, C) [or batch-first]; reference_points
        (bs, num_query, num_levels, 2 or 4); returns the shape of `query`."""
        if 'residual' in kwargs:  # deprecated_api_warning({'residual': 'identity'})
            warnings.warn('"residual" is deprecated in `MultiScaleDeformableAttention`, '
                          'please use "identity" instead')
            identity = kwargs.pop('residual')
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)

        bs, num_query, _ = query.shape
        bs, num_value, _ = value.shape
        # the reference asserts sum(H*W) == num_value with a device->host sync
        # per call (multi_scale_deform_attn.py:367); the kernel cannot read
        # outside `value` for in-range levels, so the sync is not reproduced.

        value = self._project(self.value_proj, value, key_padding_mask, 1, self.value_dtype)
        value = value.view(bs, num_value, self.num_heads, -1)
        sampling_offsets = self._project(self.sampling_offsets, query).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        attention_weights = self._project(self.attention_weights, query).view(
            bs, num_query, self.num_heads, self.num_levels * self.num_points)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(f'Last dim of reference_points must be 2 or 4, '
                             f'but get {reference_points.shape[-1]} instead.')
        if self._can_fuse(value, sampling_offsets):
            if reference_points.shape[-1] == 2:
                ref, scale = reference_points.unsqueeze(3), None        # off / (W_l, H_l) in-kernel
            else:
                ref = reference_points[..., :2].unsqueeze(3)
                scale = reference_points[..., 2:] * (0.5 / self.num_points)
            output = _run_fused(value, spatial_shapes, level_start_index, sampling_offsets,
                                attention_weights, ref, scale)
            return self._project_out(output, identity, not self.batch_first)
        attention_weights = attention_weights.softmax(-1).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.num_points \
                * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f'Last dim of reference_points must be 2 or 4, '
                             f'but get {reference_points.shape[-1]} instead.')
        output = _run_op(value, spatial_shapes, level_start_index, sampling_locations,
                         attention_weights, self.im2col_step)
        return self._project_out(output, identity, not self.batch_first)


def _fold(t, sl):
    """(G, Q, ...) -> rows sl of it as one batch entry (1, n*Q, ...); sl None: unchanged."""
    if sl is None:
        return t
    t = t[sl]
    return t.reshape((1, t.shape[0] * t.shape[1]) + tuple(t.shape[2:]))


def _pose_box_wh(kpts):
    """kpts (..., K, 2) -> (..., 1, 2): clamped extent of the pose
    (transformer.py:402-410)."""
    lo = kpts.min(dim=-2, keepdim=True)[0]
    hi = kpts.max(dim=-2, keepdim=True)[0]
    return (hi - lo).clamp(min=1e-4)


@OPERA_ATTENTION.register_module()
class MultiScaleDeformablePoseAttention(_DeformAttnBase):
    """PETR pose attention: one sampling point per keypoint, offsets scaled by
    half the pose box.  Mirrors transformer.py:251-427."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=17, im2col_step=64,
                 dropout=0.1, norm_cfg=None, init_cfg=None, batch_first=False, value_dtype=None):
        super().__init__(embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                         batch_first, norm_cfg, init_cfg, value_dtype)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        _const_linear(self.sampling_offsets, 0.)
        _const_linear(self.attention_weights, 0., 0.)
        _xavier_linear(self.value_proj)
        _xavier_linear(self.output_proj)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        """query (num_query, bs, C); value (num_key, bs, C); reference_points
        (bs, num_query, num_levels, 2K) with K == num_points."""
        if key is None:
            key = query
        if value is None:
            value = key
        # the reference only defines the residual when `residual is None`
        # (transformer.py:373-374); a given residual is honoured here
        inp_residual = query if residual is None else residual
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)

        bs, num_query, _ = query.shape
        bs, num_key, _ = value.shape

        value = self._project(self.value_proj, value, key_padding_mask, 1, self.value_dtype)
        value = value.view(bs, num_key, self.num_heads, -1)
        sampling_offsets = self._project(self.sampling_offsets, query).view(
            bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        attention_weights = self._project(self.attention_weights, query).view(
            bs, num_query, self.num_heads, self.num_levels * self.num_points)
        if reference_points.shape[-1] != self.num_points * 2:
            raise ValueError(f'Last dim of reference_points must be 2K, '
                             f'but get {reference_points.shape[-1]} instead.')
        kpts = reference_points.reshape(bs, num_query, self.num_levels, -1, 2)
        if self._can_fuse(value, sampling_offsets):
            output = _run_fused(value, spatial_shapes, level_start_index, sampling_offsets,
                                attention_weights, kpts, _pose_box_wh(kpts).squeeze(3) * 0.5)
        else:
            attention_weights = attention_weights.softmax(-1).view(
                bs, num_query, self.num_heads, self.num_levels, self.num_points)
            kpts = kpts.unsqueeze(2)
            sampling_locations = kpts + sampling_offsets * _pose_box_wh(kpts) * 0.5
            output = _run_op(value, spatial_shapes, level_start_index, sampling_locations,
                             attention_weights, self.im2col_step)
        # the reference permutes unconditionally here (transformer.py:425)
        output = self._project(self.output_proj, output).permute(1, 0, 2)
        return self.dropout(output) + inp_residual


class _MulFramesBase(_DeformAttnBase):
    """Per-frame offset / weight projections named as in the reference
    (`pre_sampling_offsets`, `sampling_offsets`, `next_sampling_offsets`, …)."""

    def _make_frame_layers(self, num_frames):
        self.num_frames = num_frames
        self._prefixes = _FRAME_PREFIXES[num_frames]
        C, M, L, P = self.embed_dims, self.num_heads, self.num_levels, self.num_points
        for pre in self._prefixes:
            setattr(self, pre + 'sampling_offsets', nn.Linear(C, M * L * P * 2))
            setattr(self, pre + 'attention_weights', nn.Linear(C, M * L * P))
        self.value_proj = nn.Linear(C, C)
        self.output_proj = nn.Linear(C, C)

    def _frame_layers(self, kind):
        return [getattr(self, pre + kind) for pre in self._prefixes]

    def _stacked_linear(self, query, kind, per_head):
        """All T projections of `kind` as one GEMM, output laid out
        (..., M, T, per_head) so it can be viewed as T*L levels directly."""
        layers = self._frame_layers(kind)
        M, T, C = self.num_heads, self.num_frames, self.embed_dims
        weight = torch.stack([l.weight.view(M, per_head, C) for l in layers], 1)
        bias = torch.stack([l.bias.view(M, per_head) for l in layers], 1)
        return F.linear(query, weight.reshape(M * T * per_head, C), bias.reshape(-1))

    def _fuse_reference_style(self, outs, logits):
        """sum_t out_t * Z_t / sum(Z), Z_t = sum exp(logits_t), no
        max-subtraction — exactly transformer.py:1736-1741,1854-1858."""
        zs = [torch.exp(lg).sum(-1, keepdim=True) for lg in logits]
        z_all = sum(zs)
        return sum(o * (z / z_all) for o, z in zip(outs, zs))


class _MulFramesPoseAttention(_MulFramesBase):
    """PAVE-Net pose-aware attention over the T frames of a clip."""

    def _init_common(self, num_frames, embed_dims, num_heads, num_levels, num_points,
                     im2col_step, dropout, norm_cfg, init_cfg, batch_first, fused, value_dtype):
        _DeformAttnBase.__init__(self, embed_dims, num_heads, num_levels, num_points, im2col_step,
                                 dropout, batch_first, norm_cfg, init_cfg, value_dtype)
        self.tag = 1  # kept: attribute of the reference class (transformer.py:1587)
        self.fused = fused
        self._make_frame_layers(num_frames)
        self.init_weights()

    def init_weights(self):
        for layer in self._frame_layers('sampling_offsets'):
            _const_linear(layer, 0.)
        for layer in self._frame_layers('attention_weights'):
            _const_linear(layer, 0., 0.)
        _xavier_linear(self.value_proj)
        _xavier_linear(self.output_proj)

    def forward(self, query, key, value, residual=None, query_pos=None, query_time_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        """query (num_query, clips, C); value (num_key, clips*T, C) with the T
        frames of a clip adjacent; key_padding_mask (clips*T, num_key);
        reference_points (clips, T*num_query, num_levels, 2K), frame-major."""
        if key is None:
            key = query
        if value is None:
            value = key
        inp_residual = query if residual is None else residual
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2)
        T, M, L, P = self.num_frames, self.num_heads, self.num_levels, self.num_points
        num_query = query.shape[1]
        bs_frames, num_key, _ = value.shape
        bs = bs_frames // T
        if reference_points.shape[-1] != P * 2:
            raise ValueError(f'Last dim of reference_points must be 2K, '
                             f'but get {reference_points.shape[-1]} instead.')
        # NOTE the order: mask first, then project (transformer.py:1706-1711),
        # unlike the single-frame classes
        value = self._project(self.value_proj, value, key_padding_mask, 2, self.value_dtype)

        # (bs, T, Q, L, K, 2): keypoints of every frame, and their pose boxes
        kpts = reference_points.reshape(bs, T, num_query, L, P, 2)
        wh = _pose_box_wh(kpts)

        if self.fused:
            offsets = self._stacked_linear(query, 'sampling_offsets', L * P * 2).view(
                bs, num_query, M, T * L, P, 2)
            logits = self._stacked_linear(query, 'attention_weights', L * P).view(
                bs, num_query, M, T * L * P)
            kpts_f = kpts.permute(0, 2, 1, 3, 4, 5).reshape(bs, num_query, T * L, P, 2)
            wh_f = wh.permute(0, 2, 1, 3, 4, 5).reshape(bs, num_query, T * L, 2)
            shapes_f, starts_f = fuse_frames_as_levels(spatial_shapes, level_start_index, T,
                                                       num_key)
            # frames of clip b are rows b*T .. b*T+T-1: a free reinterpretation
            value_f = value.reshape(bs, T * num_key, M, -1)
            if self._can_fuse(value_f, offsets):
                output = _run_fused(value_f, shapes_f, starts_f, offsets, logits, kpts_f,
                                    wh_f * 0.5)
            else:
                weights = logits.softmax(-1).view(bs, num_query, M, T * L, P)
                locations = kpts_f.unsqueeze(2) + offsets * wh_f[:, :, None, :, None, :] * 0.5
                output = _run_op(value_f, shapes_f, starts_f, locations, weights,
                                 self.im2col_step)
        else:
            outs, logits = [], []
            for t, pre in enumerate(self._prefixes):
                v_t = value[t::T].reshape(bs, num_key, M, -1).contiguous()
                off_t = getattr(self, pre + 'sampling_offsets')(query).view(
                    bs, num_query, M, L, P, 2)
                lg_t = getattr(self, pre + 'attention_weights')(query).view(
                    bs, num_query, M, L * P)
                logits.append(lg_t)
                w_t = lg_t.softmax(-1).view(bs, num_query, M, L, P)
                loc_t = kpts[:, t].unsqueeze(2) + off_t * wh[:, t].unsqueeze(2) * 0.5
                outs.append(_run_op(v_t, spatial_shapes, level_start_index, loc_t, w_t,
                                    self.im2col_step).reshape(bs, num_query, M, -1))
            output = self._fuse_reference_style(outs, logits).flatten(-2, -1)

        output = self._project(self.output_proj, output).permute(1, 0, 2)
        return self.dropout(output) + inp_residual


@OPERA_ATTENTION.register_module()
class MulFramesMultiScaleDeformablePoseAttentionNumFrames3(_MulFramesPoseAttention):
    """Mirrors transformer.py:1543-1863 (3 frames: pre / now / next)."""

    def __init__(self, num_frames=3, embed_dims=256, num_heads=8, num_levels=4, num_points=17,
                 im2col_step=64, dropout=0.1, norm_cfg=None, init_cfg=None, batch_first=False,
                 fused=True, value_dtype=None):
        if num_frames != 3:
            raise ValueError('this class has exactly 3 per-frame projections '
                             '(pre/now/next); got num_frames=%r' % (num_frames,))
        self._init_common(3, embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                          norm_cfg, init_cfg, batch_first, fused, value_dtype)


@OPERA_ATTENTION.register_module()
class MulFramesMultiScaleDeformablePoseAttentionNumFrames5(_MulFramesPoseAttention):
    """Mirrors transformer.py:2738-3114 (5 frames; no `num_frames` kwarg)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=17, im2col_step=64,
                 dropout=0.1, norm_cfg=None, init_cfg=None, batch_first=False, fused=True,
                 value_dtype=None):
        self._init_common(5, embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                          norm_cfg, init_cfg, batch_first, fused, value_dtype)


class _MulFramesJointAttention(_MulFramesBase):
    """Joint-decoder cross-attention over T frames (value is (S, G, T, C))."""

    def _init_common(self, num_frames, embed_dims, num_heads, num_levels, num_points,
                     im2col_step, dropout, batch_first, norm_cfg, init_cfg, fused, value_dtype):
        _DeformAttnBase.__init__(self, embed_dims, num_heads, num_levels, num_points, im2col_step,
                                 dropout, batch_first, norm_cfg, init_cfg, value_dtype)
        self.fused = fused
        self._make_frame_layers(num_frames)
        self.init_weights()

    def init_weights(self):
        ring = _ring_offsets(self.num_heads, self.num_levels, self.num_points)
        for layer in self._frame_layers('sampling_offsets'):
            _const_linear(layer, 0.)
            with torch.no_grad():
                layer.bias.copy_(ring)
        for layer in self._frame_layers('attention_weights'):
            _const_linear(layer, 0., 0.)
        _xavier_linear(self.value_proj)
        _xavier_linear(self.output_proj)
        self._is_init = True

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                query_time_pos=None, key_padding_mask=None, reference_points=None,
                spatial_shapes=None, level_start_index=None, value_group_sizes=None, **kwargs):
        """query (num_query, G, C); value (num_key, G, T, C); key_padding_mask
        (G, T, num_key); reference_points (T*G, num_query, num_levels, 2)
        frame-major, or (G, num_query, num_levels, 4) boxes shared by all frames.

        value_group_sizes (not in the reference): a list of ints, one per clip.  The
        reference gathers the clip's memory once per person before calling this module
        (`memory[:, img_inds]`), so value_proj and the sampling kernels see G copies of the
        same tokens.  With this argument `value` and `key_padding_mask` hold ONE entry per
        clip -- (num_key, clips, T, C) and (clips, T, num_key) -- and the G persons, ordered
        by clip, share it: person g of clip b reads value[:, b].  Same result, 1/G-th of
        the projection work and no gathered copy."""
        if 'residual' in kwargs:
            identity = kwargs.pop('residual')
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query = query.permute(1, 0, 2)
            value = value.permute(1, 0, 2, 3)
        T, M, L, P = self.num_frames, self.num_heads, self.num_levels, self.num_points
        bs, num_query, _ = query.shape
        num_value, num_frames = value.shape[1], value.shape[2]
        if num_frames != T:
            raise ValueError('value holds %d frames, module built for %d' % (num_frames, T))
        groups = None
        if value_group_sizes is not None:
            groups = [int(n) for n in value_group_sizes]
            if len(groups) != value.shape[0] or sum(groups) != bs or min(groups) < 0:
                raise ValueError('value_group_sizes %r does not split %d queries over %d clips'
                                 % (groups, bs, value.shape[0]))
            if not self.fused:
                # reference-style composition: materialise the gather the reference would have done
                idx = torch.repeat_interleave(
                    torch.arange(len(groups), device=value.device),
                    torch.tensor(groups, device=value.device), output_size=bs)
                value = value[idx]
                if key_padding_mask is not None:
                    key_padding_mask = key_padding_mask[idx]
                groups = None
        elif value.shape[0] != bs:
            raise ValueError('value holds %d batch entries, query %d' % (value.shape[0], bs))
        frame_major = value.permute(0, 2, 1, 3)              # (G, T, S, C)
        if self.fused and frame_major.is_contiguous():
            # tokens already frame-major (a batch-first encoder output viewed per clip): project in
            # place, the result IS the T*L-level value the fused call needs -- no copy either side
            value = self._project(self.value_proj, frame_major, key_padding_mask, 2,
                                  self.value_dtype).permute(0, 2, 1, 3)             # (G, S, T, C) view
        else:
            value = self._project(
                self.value_proj, value,
                None if key_padding_mask is None else key_padding_mask.transpose(1, 2),
                2, self.value_dtype)  # (G, S, T, C)

        ref_dim = reference_points.shape[-1]
        if ref_dim == 2:
            ref = reference_points.reshape(T, bs, num_query, L, 2)
            normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        elif ref_dim == 4:
            ref = reference_points
        else:
            raise ValueError(f'Last dim of reference_points must be 2 or 4, '
                             f'but get {ref_dim} instead.')

        if self.fused:
            offsets = self._stacked_linear(query, 'sampling_offsets', L * P * 2).view(
                bs, num_query, M, T * L, P, 2)
            logits = self._stacked_linear(query, 'attention_weights', L * P).view(
                bs, num_query, M, T * L * P)
            shapes_f, starts_f = fuse_frames_as_levels(spatial_shapes, level_start_index, T,
                                                       num_value)
            # frames are interleaved along dim 2 unless the tokens came frame-major: one copy
            # at most (the reference makes T `.contiguous()` copies of the same bytes)
            value_f = value.permute(0, 2, 1, 3).reshape(value.shape[0], T * num_value, M, -1)
            if ref_dim == 2:
                ref_f = ref.permute(1, 2, 0, 3, 4).reshape(bs, num_query, T * L, 1, 2)
                scale_f = None
            else:
                boxes = ref.repeat(1, 1, T, 1)
                ref_f = boxes[..., :2].unsqueeze(3)
                scale_f = boxes[..., 2:] * (0.5 / P)
            if self._can_fuse(value_f, offsets):
                def sample(v, sl):
                    return _run_fused(v, shapes_f, starts_f, _fold(offsets, sl), _fold(logits, sl),
                                      _fold(ref_f, sl), None if scale_f is None else _fold(scale_f, sl))
            else:
                weights = logits.softmax(-1).view(bs, num_query, M, T * L, P)
                if ref_dim == 2:
                    locations = ref_f.unsqueeze(2) \
                        + offsets / normalizer.repeat(T, 1)[None, None, None, :, None, :]
                else:
                    locations = ref_f.unsqueeze(2) \
                        + offsets / P * boxes[:, :, None, :, None, 2:] * 0.5

                def sample(v, sl):
                    return _run_op(v, shapes_f, starts_f, _fold(locations, sl), _fold(weights, sl),
                                   self.im2col_step)
            if groups is None:
                output = sample(value_f, None)
            else:
                # the persons of one clip become extra queries of a single batch entry
                outs, g0 = [], 0
                for b, n in enumerate(groups):
                    if n:
                        outs.append(sample(value_f[b:b + 1], slice(g0, g0 + n)).view(n, num_query, -1))
                    g0 += n
                output = outs[0] if len(outs) == 1 else torch.cat(outs)
        else:
            outs, logits = [], []
            for t, pre in enumerate(self._prefixes):
                v_t = value[:, :, t].reshape(bs, num_value, M, -1).contiguous()
                off_t = getattr(self, pre + 'sampling_offsets')(query).view(
                    bs, num_query, M, L, P, 2)
                lg_t = getattr(self, pre + 'attention_weights')(query).view(
                    bs, num_query, M, L * P)
                logits.append(lg_t)
                w_t = lg_t.softmax(-1).view(bs, num_query, M, L, P)
                if ref_dim == 2:
                    loc_t = ref[t][:, :, None, :, None, :] \
                        + off_t / normalizer[None, None, None, :, None, :]
                else:
                    loc_t = ref[:, :, None, :, None, :2] \
                        + off_t / P * ref[:, :, None, :, None, 2:] * 0.5
                outs.append(_run_op(v_t, spatial_shapes, level_start_index, loc_t, w_t,
                                    self.im2col_step).reshape(bs, num_query, M, -1))
            output = self._fuse_reference_style(outs, logits).flatten(-2, -1)

        return self._project_out(output, identity, not self.batch_first)


@ATTENTION.register_module()
class MulFramesMultiScaleDeformableAttentionNumFrames3(_MulFramesJointAttention):
    """Mirrors multi_scale_deform_attn.py:1268-1587."""

    def __init__(self, num_frames=3, embed_dims=256, num_heads=8, num_levels=4, num_points=4,
                 im2col_step=64, dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None,
                 fused=True, value_dtype=None):
        if num_frames != 3:
            raise ValueError('this class has exactly 3 per-frame projections '
                             '(pre/now/next); got num_frames=%r' % (num_frames,))
        self._init_common(3, embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                          batch_first, norm_cfg, init_cfg, fused, value_dtype)


@ATTENTION.register_module()
class MulFramesMultiScaleDeformableAttentionNumFrames5(_MulFramesJointAttention):
    """Mirrors multi_scale_deform_attn.py:1590-1982 (no `num_frames` kwarg)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64,
                 dropout=0.1, batch_first=False, norm_cfg=None, init_cfg=None, fused=True,
                 value_dtype=None):
        self._init_common(5, embed_dims, num_heads, num_levels, num_points, im2col_step, dropout,
                          batch_first, norm_cfg, init_cfg, fused, value_dtype)


MMCV_SCOPE_CLASSES = (MultiScaleDeformableAttention,
                      MulFramesMultiScaleDeformableAttentionNumFrames3,
                      MulFramesMultiScaleDeformableAttentionNumFrames5)
OPERA_SCOPE_CLASSES = (MultiScaleDeformablePoseAttention,
                       MulFramesMultiScaleDeformablePoseAttentionNumFrames3,
                       MulFramesMultiScaleDeformablePoseAttentionNumFrames5)


@FEEDFORWARD_NETWORK.register_module()
class FFN(nn.Module):
    """Mirrors mmcv/cnn/bricks/transformer.py:1046-1120 (same constructor kwargs,
    `layers.0.0` / `layers.1` parameter names and `forward(x, identity=None)`): the
    feed-forward block that follows every attention module of the reference's
    transformer layers.  With two fcs, ReLU and 128/256/1024-wide layers on CUDA it
    runs as `fused_ffn` (tensor-core GEMMs with bias, ReLU, dropout and the residual
    in their epilogues); anything else is the op-by-op composition."""

    #: False forces the op-by-op composition (nn.Linear / cuBLAS fp32)
    tensor_core_linear = True

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__()
        if num_fcs < 2:
            raise AssertionError(f'num_fcs should be no less than 2. got {num_fcs}.')
        act_type = (act_cfg or {}).get('type', 'ReLU')
        acts = {'ReLU': nn.ReLU, 'GELU': nn.GELU, 'LeakyReLU': nn.LeakyReLU, 'Tanh': nn.Tanh,
                'Sigmoid': nn.Sigmoid, 'PReLU': nn.PReLU, 'ELU': nn.ELU, 'ReLU6': nn.ReLU6}
        if act_type not in acts:
            raise KeyError('%s is not in the activation layer registry' % act_type)
        act_kwargs = {k: v for k, v in (act_cfg or {}).items() if k != 'type'}
        if act_type in ('GELU', 'Tanh', 'Sigmoid', 'PReLU'):
            act_kwargs.pop('inplace', None)
        self.embed_dims, self.feedforward_channels = embed_dims, feedforward_channels
        self.num_fcs, self.act_cfg = num_fcs, act_cfg
        self.activate = acts[act_type](**act_kwargs)
        layers, in_channels = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_channels), self.activate,
                                        nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        if dropout_layer:
            kind = dropout_layer.get('type', 'Dropout')
            if kind != 'Dropout':
                raise KeyError('dropout_layer type %s is not supported here (Dropout only)' % kind)
            self.dropout_layer = nn.Dropout(dropout_layer.get('drop_prob', 0.5))
        else:
            self.dropout_layer = nn.Identity()
        self.add_identity = add_identity
        self.ffn_drop = float(ffn_drop)

    def _fusable(self, x):
        return (self.tensor_core_linear and self.num_fcs == 2 and isinstance(self.activate, nn.ReLU)
                and isinstance(self.dropout_layer, nn.Identity)
                and ffn_supported(x, self.layers[0][0].weight, self.layers[1].weight))

    def forward(self, x, identity=None):
        if not x.is_cuda:
            raise RuntimeError('pavenet_b200 modules run on CUDA tensors only (there is no CPU '
                               'fallback); got x on %s' % x.device)
        if self._fusable(x):
            fc1, fc2 = self.layers[0][0], self.layers[1]
            return fused_ffn(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias,
                             self.ffn_drop if self.training else 0.0, identity, self.add_identity)
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm (what mmcv's `build_norm_layer(dict(type='LN'), 256)` returns for the `norm`
    steps of the transformer layers; same parameters and state-dict names) running on the
    streaming kernels of layernorm.cu when the input is fp32 CUDA with 256 channels."""

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('pavenet_b200 modules run on CUDA tensors only (there is no CPU '
                               'fallback); got x on %s' % x.device)
        if tuple(self.normalized_shape) == (256,) and layer_norm_supported(x, self.weight, self.bias):
            return layer_norm256(x, self.weight, self.bias, self.eps)
        return super().forward(x)
