"""Operator boundary: `MultiScaleDeformableAttnFunction` and the `ext_module`
shim, with the reference's exact names, argument order and error behaviour,
running on the sm_100a kernels behind the C ABI (`include/pavenet_msda.h`).

Reference being mirrored:
  * `MultiScaleDeformableAttnFunction`  third_party/mmcv/mmcv/ops/multi_scale_deform_attn.py:20-89
  * `ext_module.ms_deform_attn_forward/backward` (pybind kwargs)
        third_party/mmcv/mmcv/ops/csrc/pytorch/pybind.cpp:737-748
  * host-side checks  third_party/mmcv/mmcv/ops/csrc/pytorch/cuda/ms_deform_attn_cuda.cu:209-245, 279-318

There is deliberately no CPU / PyTorch fallback in this module.
"""
import torch
from torch.autograd.function import Function, once_differentiable

from . import _capi

__all__ = [
    'MultiScaleDeformableAttnFunction', 'ext_module', 'ms_deform_attn_forward',
    'ms_deform_attn_backward', 'fuse_frames_as_levels', 'BF16_GRAD_VALUE_ATOMICS',
    'EARLY_GRAD_VALUE_CLEAR',
    'HostWorkspace', 'FusedMultiScaleDeformableAttnFunction', 'fused_supported',
    'Linear256Function', 'linear256', 'linear256_supported',
    'FusedFFNFunction', 'fused_ffn', 'ffn_supported', 'device_dropout_seed',
    'LayerNorm256Function', 'layer_norm256', 'layer_norm_supported',
]

#: When value is stored in bf16, accumulate grad_value in an fp32 scratch
#: buffer and round once at the end (False, default, accurate) or let the
#: kernel reduce straight into a bf16 tensor with packed bf16x2 atomics
#: (True: half the scatter bytes, every partial sum rounded to 8 bits).
BF16_GRAD_VALUE_ATOMICS = False

#: Zero-fill the backward's grad_value buffer during the FORWARD call (when `value`
#: requires grad): for the small-Q shapes the fill is folded into the persistent forward
#: kernel (msda_forward_clear), which takes the full-tensor clear off the critical path of
#: the backward.  The buffer then lives from forward to backward.
EARLY_GRAD_VALUE_CLEAR = True

_DTYPE_CODE = {
    torch.float32: _capi.MSDA_F32,
    torch.float64: _capi.MSDA_F64,
    torch.bfloat16: _capi.MSDA_BF16,
}


def _check_inputs(value, spatial_shapes, level_start_index, sampling_loc,
                  attn_weight, im2col_step):
    """The reference's AT_ASSERTM / device-consistency checks, as RuntimeError."""
    named = (('value', value), ('spatial_shapes', spatial_shapes),
             ('level_start_index', level_start_index),
             ('sampling_loc', sampling_loc), ('attn_weight', attn_weight))
    for name, t in named:
        if not isinstance(t, torch.Tensor):
            raise TypeError('%s must be a torch.Tensor, got %s' % (name, type(t)))
        if not t.is_contiguous():
            raise RuntimeError('%s tensor has to be contiguous' % name)
        if not t.is_cuda:
            raise RuntimeError('%s must be a CUDA tensor' % name)
        if t.device != value.device:
            # pytorch_device_registry.hpp:116-122
            raise RuntimeError('%s is on %s but value is on %s: all tensors must '
                               'be on the same device' % (name, t.device, value.device))
    if value.dim() != 4:
        raise RuntimeError('value must have shape (bs, num_keys, num_heads, dim_per_head), '
                           'got %s' % (tuple(value.shape),))
    if sampling_loc.dim() != 6 or sampling_loc.shape[-1] != 2:
        raise RuntimeError('sampling_locations must have shape (bs, num_queries, num_heads, '
                           'num_levels, num_points, 2), got %s' % (tuple(sampling_loc.shape),))
    B, S, M, D = value.shape
    Bq, Q, Mq, L, P, _ = sampling_loc.shape
    if (Bq, Mq) != (B, M):
        raise RuntimeError('sampling_locations %s does not match value %s in batch / heads'
                           % (tuple(sampling_loc.shape), tuple(value.shape)))
    if tuple(attn_weight.shape) != (B, Q, M, L, P):
        raise RuntimeError('attention_weights must have shape %s, got %s'
                           % ((B, Q, M, L, P), tuple(attn_weight.shape)))
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError('spatial_shapes and level_start_index must be int64 tensors')
    if tuple(spatial_shapes.shape) != (L, 2) or tuple(level_start_index.shape) != (L,):
        raise RuntimeError('spatial_shapes must be (%d, 2) and level_start_index (%d,), got %s / %s'
                           % (L, L, tuple(spatial_shapes.shape), tuple(level_start_index.shape)))
    if sampling_loc.dtype not in (torch.float32, torch.float64):
        raise RuntimeError('sampling_locations must be float32 or float64, got %s'
                           % sampling_loc.dtype)
    if attn_weight.dtype != sampling_loc.dtype:
        raise RuntimeError('attention_weights dtype %s != sampling_locations dtype %s'
                           % (attn_weight.dtype, sampling_loc.dtype))
    if value.dtype != sampling_loc.dtype and not (
            value.dtype == torch.bfloat16 and sampling_loc.dtype == torch.float32):
        raise RuntimeError('value dtype %s incompatible with sampling_locations dtype %s '
                           '(same dtype, or bfloat16 value with float32 locations)'
                           % (value.dtype, sampling_loc.dtype))
    step = min(B, int(im2col_step))
    if B > 0 and (step <= 0 or B % step != 0):
        # ms_deform_attn_cuda.cu:242-245
        raise RuntimeError('batch(%d) must divide im2col_step(%d)' % (B, step))
    return B, S, M, D, L, Q, P


def _check_clear(clear, device):
    if clear is None:
        return None, 0
    if not (isinstance(clear, torch.Tensor) and clear.is_cuda and clear.device == device
            and clear.is_contiguous()):
        raise RuntimeError('clear must be a contiguous CUDA tensor on %s' % device)
    nbytes = clear.numel() * clear.element_size()
    return (clear.data_ptr(), nbytes) if nbytes else (None, 0)


def ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                           sampling_locations, attention_weights, im2col_step=64, clear=None):
    """Drop-in for `ext_module.ms_deform_attn_forward` (pybind.cpp:737-742).

    Returns a new tensor of shape (bs, num_queries, num_heads*dim_per_head)
    in the dtype of `sampling_locations`.  `clear` (not in the reference): a tensor to
    zero-fill on the same stream, normally the coming backward's grad_value
    (`msda_forward_clear`).
    """
    B, S, M, D, L, Q, P = _check_inputs(value, value_spatial_shapes, value_level_start_index,
                                        sampling_locations, attention_weights, im2col_step)
    lib = _capi.load()
    if min(B, S, M, D, L, Q, P) == 0:
        # nothing to sample (or nothing to write): the sum over an empty set
        if clear is not None:
            clear.zero_()
        return torch.zeros((B, Q, M * D), dtype=sampling_locations.dtype, device=value.device)
    clear_ptr, clear_bytes = _check_clear(clear, value.device)
    with torch.cuda.device(value.device):
        output = torch.empty((B, Q, M * D), dtype=sampling_locations.dtype, device=value.device)
        stream = torch.cuda.current_stream().cuda_stream
        status = lib.msda_forward_clear(
            value.data_ptr(), value_spatial_shapes.data_ptr(), value_level_start_index.data_ptr(),
            sampling_locations.data_ptr(), attention_weights.data_ptr(), output.data_ptr(),
            B, S, M, D, L, Q, P, _DTYPE_CODE[sampling_locations.dtype], _DTYPE_CODE[value.dtype],
            clear_ptr, clear_bytes, stream)
    _capi.check(status, 'msda_forward')
    return output


def ms_deform_attn_backward(value, value_spatial_shapes, value_level_start_index,
                            sampling_locations, attention_weights, grad_output, grad_value,
                            grad_sampling_loc, grad_attn_weight, im2col_step=64):
    """Drop-in for `ext_module.ms_deform_attn_backward` (pybind.cpp:743-748).

    Accumulates into `grad_value` (caller zero-fills it, as
    multi_scale_deform_attn.py:72-74 does) and overwrites the other two.
    """
    B, S, M, D, L, Q, P = _check_inputs(value, value_spatial_shapes, value_level_start_index,
                                        sampling_locations, attention_weights, im2col_step)
    for name, t, like in (('grad_output', grad_output, None),
                          ('grad_value', grad_value, value),
                          ('grad_sampling_loc', grad_sampling_loc, sampling_locations),
                          ('grad_attn_weight', grad_attn_weight, attention_weights)):
        if not t.is_contiguous():
            raise RuntimeError('%s tensor has to be contiguous' % name)
        if not t.is_cuda or t.device != value.device:
            raise RuntimeError('%s must be a CUDA tensor on %s' % (name, value.device))
        if like is not None and t.shape != like.shape:
            raise RuntimeError('%s has shape %s, expected %s'
                               % (name, tuple(t.shape), tuple(like.shape)))
    if grad_output.numel() != B * Q * M * D or grad_output.dtype != sampling_locations.dtype:
        raise RuntimeError('grad_output must hold %d %s elements, got %s %s'
                           % (B * Q * M * D, sampling_locations.dtype,
                              tuple(grad_output.shape), grad_output.dtype))
    if (grad_sampling_loc.dtype != sampling_locations.dtype or
            grad_attn_weight.dtype != sampling_locations.dtype):
        raise RuntimeError('location / weight gradients must have dtype %s'
                           % sampling_locations.dtype)
    if grad_value.dtype not in (sampling_locations.dtype, value.dtype):
        raise RuntimeError('grad_value dtype %s must be %s or %s'
                           % (grad_value.dtype, sampling_locations.dtype, value.dtype))
    lib = _capi.load()
    if min(B, S, M, D, L, Q, P) == 0:
        grad_sampling_loc.zero_()
        grad_attn_weight.zero_()
        return
    with torch.cuda.device(value.device):
        stream = torch.cuda.current_stream().cuda_stream
        status = lib.msda_backward(
            value.data_ptr(), value_spatial_shapes.data_ptr(), value_level_start_index.data_ptr(),
            sampling_locations.data_ptr(), attention_weights.data_ptr(), grad_output.data_ptr(),
            grad_value.data_ptr(), grad_sampling_loc.data_ptr(), grad_attn_weight.data_ptr(),
            B, S, M, D, L, Q, P, _DTYPE_CODE[sampling_locations.dtype], _DTYPE_CODE[value.dtype],
            _DTYPE_CODE[grad_value.dtype], stream)
    _capi.check(status, 'msda_backward')


class _ExtModule(object):
    """Stands in for `mmcv._ext` as far as this op is concerned
    (multi_scale_deform_attn.py:16-17 loads exactly these two attributes)."""
    ms_deform_attn_forward = staticmethod(ms_deform_attn_forward)
    ms_deform_attn_backward = staticmethod(ms_deform_attn_backward)


ext_module = _ExtModule()


def _grad_value_dtype(value, sampling_locations):
    if value.dtype == torch.bfloat16 and BF16_GRAD_VALUE_ATOMICS:
        return torch.bfloat16
    return sampling_locations.dtype


class MultiScaleDeformableAttnFunction(Function):
    """Same `apply` signature and return shape as the reference
    (multi_scale_deform_attn.py:20-89)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        """
        Args:
            value: (bs, num_keys, num_heads, embed_dims // num_heads)
            value_spatial_shapes: (num_levels, 2) int64 on the GPU, (h, w)
            value_level_start_index: (num_levels,) int64 on the GPU
            sampling_locations: (bs, num_queries, num_heads, num_levels, num_points, 2),
                normalised (x, y)
            attention_weights: (bs, num_queries, num_heads, num_levels, num_points)
            im2col_step: int; only validated (batch % min(batch, step) == 0),
                it never changes results

        Returns:
            (bs, num_queries, embed_dims)
        """
        ctx.im2col_step = im2col_step
        ctx.grad_value_buf = None
        if (EARLY_GRAD_VALUE_CLEAR and ctx.needs_input_grad[0] and value.is_cuda
                and value.numel() > 0):
            # the backward's accumulator, zero-filled by the forward launch
            ctx.grad_value_buf = torch.empty(value.shape, dtype=_grad_value_dtype(
                value, sampling_locations), device=value.device)
            output = ms_deform_attn_forward(
                value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step=ctx.im2col_step, clear=ctx.grad_value_buf)
        else:
            output = ext_module.ms_deform_attn_forward(
                value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step=ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, \
            sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, ctx.grad_value_buf = ctx.grad_value_buf, None   # used once
        if grad_value is None:
            grad_value = torch.zeros(value.shape, dtype=_grad_value_dtype(value, sampling_locations),
                                     device=value.device)
        # fully overwritten by the kernels: no zero-fill needed
        grad_sampling_loc = torch.empty_like(sampling_locations)
        grad_attn_weight = torch.empty_like(attention_weights)
        ext_module.ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations,
            attention_weights, grad_output.contiguous(), grad_value, grad_sampling_loc,
            grad_attn_weight, im2col_step=ctx.im2col_step)
        if grad_value.dtype != value.dtype:
            grad_value = grad_value.to(value.dtype)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


def fuse_frames_as_levels(spatial_shapes, level_start_index, num_frames, num_keys):
    """Level tables for sampling T frames in ONE op call.

    The T frames of a clip are adjacent in the batch dimension, so
    `value.view(clips, T*num_keys, heads, dim)` is a zero-copy view in which
    frame t's level l is "level" t*L+l, starting at `t*num_keys + start_l`
    (SURVEY.md section 3.3).  Built with device ops only: no host sync.
    """
    shapes = spatial_shapes.repeat(num_frames, 1)
    frame_base = torch.arange(num_frames, device=level_start_index.device,
                              dtype=level_start_index.dtype) * num_keys
    starts = (frame_base[:, None] + level_start_index[None, :]).reshape(-1)
    return shapes.contiguous(), starts.contiguous()


class HostWorkspace(object):
    """Host-buffer front-end of the C ABI (`msda_forward_host`,
    `msda_forward_backward_host`): CPU tensors in, CPU tensors out.

    The staged call is pipelined inside the library (upload of the next piece,
    kernels of the current one and download of the previous one overlap on
    three streams), so pass PINNED tensors (`tensor.pin_memory()`) to get the
    overlap; pageable tensors work but serialise.  This is the entry point a
    caller without device-memory management binds;
    the autograd `MultiScaleDeformableAttnFunction` is the one PyTorch uses.
    """

    def __init__(self):
        import ctypes
        self._lib = _capi.load()
        handle = ctypes.c_void_p()
        _capi.check(self._lib.msda_workspace_create(ctypes.byref(handle)), 'msda_workspace_create')
        self._ws = handle
        self._keepalive = None

    def set_piece_bytes(self, nbytes):
        """Upload bytes per pipeline piece (default 12 MiB)."""
        _capi.check(self._lib.msda_workspace_set_piece_bytes(self._ws, int(nbytes)),
                    'msda_workspace_set_piece_bytes')

    def close(self):
        if getattr(self, '_ws', None):
            self._lib.msda_workspace_destroy(self._ws)
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    @staticmethod
    def _check_host(value, shapes, lsi, loc, aw):
        for name, t in (('value', value), ('spatial_shapes', shapes), ('level_start_index', lsi),
                        ('sampling_locations', loc), ('attention_weights', aw)):
            if t.is_cuda:
                raise RuntimeError('%s must be a CPU tensor for the host-buffer entry points' % name)
            if not t.is_contiguous():
                raise RuntimeError('%s tensor has to be contiguous' % name)
        if shapes.dtype != torch.int64 or lsi.dtype != torch.int64:
            raise RuntimeError('spatial_shapes and level_start_index must be int64 tensors')
        if value.dim() != 4 or loc.dim() != 6 or loc.shape[-1] != 2:
            raise RuntimeError('value must be (bs, keys, heads, dim) and sampling_locations '
                               '(bs, queries, heads, levels, points, 2), got %s / %s'
                               % (tuple(value.shape), tuple(loc.shape)))
        B, S, M, D = value.shape
        Bq, Q, Mq, L, P, _ = loc.shape
        if min(B, S, M, D, L, Q, P) <= 0:
            raise RuntimeError('host-buffer entry points need non-empty tensors')
        if (Bq, Mq) != (B, M) or tuple(aw.shape) != (B, Q, M, L, P):
            raise RuntimeError('inconsistent shapes: value %s sampling_locations %s attention_weights %s'
                               % (tuple(value.shape), tuple(loc.shape), tuple(aw.shape)))
        if tuple(shapes.shape) != (L, 2) or tuple(lsi.shape) != (L,):
            raise RuntimeError('spatial_shapes must be (%d, 2) and level_start_index (%d,)' % (L, L))
        if loc.dtype not in (torch.float32, torch.float64) or aw.dtype != loc.dtype:
            raise RuntimeError('sampling_locations / attention_weights must share float32 or float64')
        if value.dtype != loc.dtype and not (value.dtype == torch.bfloat16
                                             and loc.dtype == torch.float32):
            raise RuntimeError('value dtype %s incompatible with %s' % (value.dtype, loc.dtype))
        return B, S, M, D, L, Q, P

    @staticmethod
    def _check_result(name, t, shape, dtype):
        """A caller-supplied result / gradient buffer the library will write with async copies."""
        if not isinstance(t, torch.Tensor) or t.is_cuda:
            raise RuntimeError('%s must be a CPU tensor' % name)
        if not t.is_contiguous():
            raise RuntimeError('%s tensor has to be contiguous' % name)
        if tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            raise RuntimeError('%s must be %s %s, got %s %s'
                               % (name, tuple(shape), dtype, tuple(t.shape), t.dtype))

    def forward(self, value, spatial_shapes, level_start_index, sampling_locations,
                attention_weights, out=None):
        B, S, M, D, L, Q, P = self._check_host(value, spatial_shapes, level_start_index,
                                               sampling_locations, attention_weights)
        if out is None:
            out = torch.empty((B, Q, M * D), dtype=sampling_locations.dtype)
        self._check_result('out', out, (B, Q, M * D), sampling_locations.dtype)
        status = self._lib.msda_forward_host(
            self._ws, value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_locations.data_ptr(), attention_weights.data_ptr(), out.data_ptr(),
            B, S, M, D, L, Q, P, _DTYPE_CODE[sampling_locations.dtype], _DTYPE_CODE[value.dtype])
        _capi.check(status, 'msda_forward_host')
        return out

    def wait(self):
        """Complete the call queued with `forward_backward(..., wait=False)`."""
        _capi.check(self._lib.msda_workspace_wait(self._ws), 'msda_workspace_wait')
        self._keepalive = None

    def forward_backward(self, value, spatial_shapes, level_start_index, sampling_locations,
                         attention_weights, grad_output, out=None, grad_value=None,
                         grad_sampling_loc=None, grad_attn_weight=None, wait=True):
        """Returns (out, grad_value, grad_sampling_loc, grad_attn_weight); the
        optional arguments are preallocated (pinned) result buffers.

        wait=False queues the call (`msda_forward_backward_host_async`) and returns at once: the
        result tensors are valid, and the inputs may be modified, only after `wait()`.  Alternating
        between two workspaces keeps the link busy across calls (the next call's first upload runs
        under this call's last download)."""
        B, S, M, D, L, Q, P = self._check_host(value, spatial_shapes, level_start_index,
                                               sampling_locations, attention_weights)
        dt = sampling_locations.dtype
        if out is None:
            out = torch.empty((B, Q, M * D), dtype=dt)
        if grad_value is None:
            grad_value = torch.empty(value.shape, dtype=dt)
        if grad_sampling_loc is None:
            grad_sampling_loc = torch.empty_like(sampling_locations)
        if grad_attn_weight is None:
            grad_attn_weight = torch.empty_like(attention_weights)
        if grad_value.dtype != dt:
            raise RuntimeError('grad_value is returned in %s' % dt)
        self._check_result('out', out, (B, Q, M * D), dt)
        self._check_result('grad_value', grad_value, value.shape, dt)
        self._check_result('grad_sampling_loc', grad_sampling_loc, sampling_locations.shape, dt)
        self._check_result('grad_attn_weight', grad_attn_weight, attention_weights.shape, dt)
        if grad_output.is_cuda or grad_output.dtype != dt or grad_output.numel() != B * Q * M * D:
            raise RuntimeError('grad_output must be a CPU %s tensor of %d elements' % (dt, B * Q * M * D))
        grad_output = grad_output.contiguous()
        entry = self._lib.msda_forward_backward_host if wait else self._lib.msda_forward_backward_host_async
        if not wait:   # keep every buffer of the queued call alive until wait()
            self._keepalive = (value, spatial_shapes, level_start_index, sampling_locations,
                               attention_weights, grad_output, out, grad_value, grad_sampling_loc,
                               grad_attn_weight)
        status = entry(
            self._ws, value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_locations.data_ptr(), attention_weights.data_ptr(), grad_output.data_ptr(),
            out.data_ptr(), grad_value.data_ptr(), grad_sampling_loc.data_ptr(),
            grad_attn_weight.data_ptr(), B, S, M, D, L, Q, P, _DTYPE_CODE[dt],
            _DTYPE_CODE[value.dtype])
        _capi.check(status, 'msda_forward_backward_host')
        return out, grad_value, grad_sampling_loc, grad_attn_weight


# ---------------------------------------------------------------------------
# fused prologue / epilogue
# ---------------------------------------------------------------------------
def fused_supported(value, offsets):
    """True when `FusedMultiScaleDeformableAttnFunction` has a kernel for these
    tensors (CUDA, fp32 projections, fp32/bf16 value, 16 / 32 / 64 channels per head)."""
    return (value.is_cuda and offsets.is_cuda and value.dim() == 4 and value.shape[-1] in (16, 32, 64)
            and offsets.dtype == torch.float32
            and value.dtype in (torch.float32, torch.bfloat16)
            and offsets.dim() == 6 and 0 < offsets.shape[3] <= 64
            and min(value.shape) > 0 and min(offsets.shape) > 0)


def _check_fused_tables(value, spatial_shapes, level_start_index, logits, ref_points, scale, L):
    """The checks `_check_inputs` makes for the plain op, for the fused entry points: the
    kernels read the level tables, logits and reference points through raw pointers, so a
    CPU / int32 / strided / other-device tensor must be refused here
    (ms_deform_attn_cuda.cu:218-232 AT_ASSERTM contiguous / is_cuda)."""
    named = [('spatial_shapes', spatial_shapes), ('level_start_index', level_start_index),
             ('logits', logits), ('ref_points', ref_points)]
    if scale is not None:
        named.append(('scale', scale))
    for name, t in named:
        if not isinstance(t, torch.Tensor):
            raise TypeError('%s must be a torch.Tensor, got %s' % (name, type(t)))
        if not t.is_cuda:
            raise RuntimeError('%s must be a CUDA tensor' % name)
        if t.device != value.device:
            raise RuntimeError('%s is on %s but value is on %s: all tensors must be on the '
                               'same device' % (name, t.device, value.device))
    for name, t in named[:2]:
        if t.dtype != torch.int64:
            raise RuntimeError('spatial_shapes and level_start_index must be int64 tensors')
        if not t.is_contiguous():
            raise RuntimeError('%s tensor has to be contiguous' % name)
    if tuple(spatial_shapes.shape) != (L, 2) or tuple(level_start_index.shape) != (L,):
        raise RuntimeError('spatial_shapes must be (%d, 2) and level_start_index (%d,), got %s / %s'
                           % (L, L, tuple(spatial_shapes.shape), tuple(level_start_index.shape)))
    if logits.dtype != torch.float32:
        raise RuntimeError('logits must be float32, got %s' % logits.dtype)
    if not ref_points.is_floating_point() or (scale is not None and not scale.is_floating_point()):
        raise RuntimeError('ref_points / scale must be floating-point tensors')


class FusedMultiScaleDeformableAttnFunction(Function):
    """The op with the modules' elementwise chain folded in.

    Replaces, in one forward and one backward launch,
        weights   = softmax(logits over L*P)
        locations = ref_points + offsets * scale        (scale None: / (W_l, H_l))
        out       = MultiScaleDeformableAttnFunction(value, ..., locations, weights)
    (multi_scale_deform_attn.py:373-401; transformer.py:390-420) so that
    `sampling_locations` / `attention_weights` and their gradients are never
    written to or re-read from HBM.

    Args:
        value (bs, num_keys, heads, D), D in {16, 32, 64}, fp32 or bf16
        spatial_shapes (L, 2) int64, level_start_index (L,) int64, on the GPU
        offsets (bs, Q, heads, L, P, 2) fp32 — raw `sampling_offsets` output
        logits (bs, Q, heads, L*P) fp32 — raw `attention_weights` output
        ref_points (bs, Q, L, R, 2) fp32, R = 1 (one reference per level) or P (one per point)
        scale (bs, Q, L, 2) fp32 or None
    Returns (bs, Q, heads*D).
    """

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, offsets, logits, ref_points, scale):
        B, S, M, D = value.shape
        _, Q, _, L, P, _ = offsets.shape
        R = ref_points.shape[3]
        if not fused_supported(value, offsets):
            raise RuntimeError('fused deformable attention: unsupported tensors '
                               '(need CUDA, fp32 projections, 16 / 32 / 64 channels per head)')
        if tuple(logits.shape) != (B, Q, M, L * P) or tuple(ref_points.shape) != (B, Q, L, R, 2) \
                or R not in (1, P) or (scale is not None and tuple(scale.shape) != (B, Q, L, 2)):
            raise RuntimeError('fused deformable attention: inconsistent shapes value %s offsets %s '
                               'logits %s ref_points %s' % (tuple(value.shape), tuple(offsets.shape),
                                                            tuple(logits.shape), tuple(ref_points.shape)))
        _check_fused_tables(value, spatial_shapes, level_start_index, logits, ref_points, scale, L)
        value, offsets, logits = value.contiguous(), offsets.contiguous(), logits.contiguous()
        ref_points = ref_points.contiguous().float()
        scale = None if scale is None else scale.contiguous().float()
        lib = _capi.load()
        ctx.grad_value_buf = None
        with torch.cuda.device(value.device):
            out = torch.empty((B, Q, M * D), dtype=torch.float32, device=value.device)
            stats = torch.empty((B, Q, M, 2), dtype=torch.float32, device=value.device)
            if EARLY_GRAD_VALUE_CLEAR and ctx.needs_input_grad[0]:
                ctx.grad_value_buf = torch.empty(value.shape, dtype=torch.float32,
                                                 device=value.device)
            clear_ptr, clear_bytes = _check_clear(ctx.grad_value_buf, value.device)
            status = lib.msda_fused_forward(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                offsets.data_ptr(), logits.data_ptr(), ref_points.data_ptr(),
                None if scale is None else scale.data_ptr(), out.data_ptr(), stats.data_ptr(),
                B, S, M, D, L, Q, P, R, _DTYPE_CODE[value.dtype], clear_ptr, clear_bytes,
                torch.cuda.current_stream().cuda_stream)
        _capi.check(status, 'msda_fused_forward')
        ctx.save_for_backward(value, spatial_shapes, level_start_index, offsets, logits,
                              ref_points, scale, stats, out)
        ctx.has_scale = scale is not None
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, spatial_shapes, level_start_index, offsets, logits, ref_points, scale, stats, out = \
            ctx.saved_tensors
        B, S, M, D = value.shape
        _, Q, _, L, P, _ = offsets.shape
        R = ref_points.shape[3]
        need_ref = ctx.needs_input_grad[5]
        need_scale = ctx.has_scale and ctx.needs_input_grad[6]
        lib = _capi.load()
        with torch.cuda.device(value.device):
            grad_value, ctx.grad_value_buf = ctx.grad_value_buf, None   # used once
            if grad_value is None:
                grad_value = torch.zeros(value.shape, dtype=torch.float32, device=value.device)
            grad_offsets = torch.empty_like(offsets)
            grad_logits = torch.empty_like(logits)
            grad_loc = torch.empty_like(offsets) if (need_ref or need_scale) else None
            status = lib.msda_fused_backward(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                offsets.data_ptr(), logits.data_ptr(), ref_points.data_ptr(),
                None if scale is None else scale.data_ptr(), stats.data_ptr(), out.data_ptr(),
                grad_output.contiguous().data_ptr(), grad_value.data_ptr(),
                grad_offsets.data_ptr(), grad_logits.data_ptr(),
                None if grad_loc is None else grad_loc.data_ptr(),
                B, S, M, D, L, Q, P, R, _DTYPE_CODE[value.dtype],
                torch.cuda.current_stream().cuda_stream)
        _capi.check(status, 'msda_fused_backward')
        if grad_value.dtype != value.dtype:
            grad_value = grad_value.to(value.dtype)
        grad_ref = grad_scale = None
        if need_ref:
            # loc = ref[b,q,l,(p)] + ...: sum over heads (and over points when R == 1)
            grad_ref = grad_loc.sum(dim=2) if R == P else grad_loc.sum(dim=(2, 4)).unsqueeze(3)
        if need_scale:
            grad_scale = (grad_loc * offsets).sum(dim=(2, 4))
        return grad_value, None, None, grad_offsets, grad_logits, grad_ref, grad_scale


# ---------------------------------------------------------------------------
# fp32 Linear layers on the tcgen05 tensor cores
# ---------------------------------------------------------------------------
_LINEAR_WIDTHS = (128, 256, 1024)


def _linear_shape_ok(n_out, n_in):
    return (n_out in _LINEAR_WIDTHS and n_in in _LINEAR_WIDTHS
            and not (n_out == n_in and n_in != 256))


def linear256_supported(x, weight):
    """True when `linear256` has kernels for these tensors: CUDA, fp32, weight
    (out, in) with both widths in {128, 256, 1024} (128x128 and 1024x1024
    excluded) -- every projection of an embed_dims = 256 attention module whose
    offsets / attention weights are 128 or 256 wide, and the 256 <-> 1024
    feed-forward pair."""
    return (x.is_cuda and weight.is_cuda and x.dtype == torch.float32
            and weight.dtype == torch.float32 and weight.dim() == 2
            and _linear_shape_ok(*weight.shape)
            and x.shape[-1] == weight.shape[1] and x.numel() > 0)


def _next_dropout_seed():
    # drawn from the CPU generator: follows torch.manual_seed, costs no GPU work or sync
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class _DeviceSeed(object):
    """Dropout seeds for code that will be captured into a CUDA graph.

    A seed passed by value is baked into the captured kernel arguments, so every replay
    would drop the same elements.  Inside `device_dropout_seed(t)` each dropout site gets a
    fixed small offset instead and the kernels add `t[0]` (a 1-element int64 CUDA tensor)
    to it on the device; refresh `t` between replays (`t.random_()`) for new masks."""
    tensor = None
    next_offset = 0


class device_dropout_seed(object):
    """`salt`: where this context's dropout sites start in seed space.  Contexts that share one
    seed tensor (the graphed stages of a step) must use different salts, or site k of every
    context would draw the same mask; `graphs.GraphedStage` passes one per captured graph."""
    _SITE_STRIDE = 0x9E3779B97F4A7C15 % (2 ** 40)      # sites far apart in seed space

    def __init__(self, seed_tensor, salt=0):
        if not (seed_tensor.is_cuda and seed_tensor.dtype == torch.int64 and seed_tensor.numel() == 1):
            raise ValueError('device_dropout_seed needs a 1-element int64 CUDA tensor')
        self.tensor = seed_tensor
        # 2^20 sites per context before two contexts could meet
        self.start = (int(salt) * (self._SITE_STRIDE << 20)) % (2 ** 62)

    def __enter__(self):
        self.saved = (_DeviceSeed.tensor, _DeviceSeed.next_offset)
        _DeviceSeed.tensor, _DeviceSeed.next_offset = self.tensor, self.start
        return self

    def __exit__(self, *exc):
        _DeviceSeed.tensor, _DeviceSeed.next_offset = self.saved
        return False


def _draw_seed():
    """-> (host seed, device seed tensor or None) for one dropout site."""
    if _DeviceSeed.tensor is not None:
        _DeviceSeed.next_offset += device_dropout_seed._SITE_STRIDE
        return _DeviceSeed.next_offset % (2 ** 62), _DeviceSeed.tensor
    return _next_dropout_seed(), None


def _linear_fused_raw(x2d, weight, bias=None, row_mask=None, mask_mode=0, relu=False, gate=None,
                      gate_scale=1.0, dropout_p=0.0, seed=0, residual=None, out_dtype=torch.float32,
                      seed_tensor=None):
    """One call of msda_linear_fused on 2-D contiguous fp32 tensors (see include/pavenet_msda.h
    for the order of the epilogue steps)."""
    lib = _capi.load()
    rows = x2d.shape[0]
    n_out, n_in = weight.shape
    with torch.cuda.device(x2d.device):
        y = torch.empty((rows, n_out), dtype=out_dtype, device=x2d.device)
        scratch = torch.empty(2 * n_out * n_in, dtype=torch.float32, device=x2d.device)
        status = lib.msda_linear_fused(
            x2d.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
            None if row_mask is None else row_mask.data_ptr(), mask_mode, int(relu),
            None if gate is None else gate.data_ptr(), gate_scale, dropout_p, seed,
            None if seed_tensor is None else seed_tensor.data_ptr(),
            None if residual is None else residual.data_ptr(), y.data_ptr(), rows,
            n_in, n_out, _DTYPE_CODE[out_dtype], scratch.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    _capi.check(status, 'msda_linear_fused')
    return y


def _wgrad_raw(g, x2d, row_mask, mask_mode, n_out, n_in):
    """dW (out, in) = g^T x2d, split-K on the tensor cores."""
    lib = _capi.load()
    with torch.cuda.device(g.device):
        grad_w = torch.empty((n_out, n_in), dtype=torch.float32, device=g.device)
        status = lib.msda_linear256_wgrad(
            g.data_ptr(), x2d.data_ptr(), None if row_mask is None else row_mask.data_ptr(),
            mask_mode, grad_w.data_ptr(), g.shape[0], n_in, n_out,
            torch.cuda.current_stream().cuda_stream)
    _capi.check(status, 'msda_linear256_wgrad')
    return grad_w


def _colsum_raw(g, row_mask=None):
    """Column sums of g (rows, width): the bias gradient, one streaming pass."""
    lib = _capi.load()
    with torch.cuda.device(g.device):
        out = torch.empty(g.shape[1], dtype=torch.float32, device=g.device)
        status = lib.msda_colsum256(g.data_ptr(), None if row_mask is None else row_mask.data_ptr(),
                                    out.data_ptr(), g.shape[0], g.shape[1],
                                    torch.cuda.current_stream().cuda_stream)
    _capi.check(status, 'msda_colsum256')
    return out


def _dropout_backward_raw(g, dropout_p, seed, want_bias, seed_tensor=None):
    """g * keep / (1 - p) with the forward's keep decisions, and its column sums."""
    lib = _capi.load()
    with torch.cuda.device(g.device):
        out = torch.empty_like(g)
        bias = torch.empty(g.shape[1], dtype=torch.float32, device=g.device) if want_bias else None
        status = lib.msda_dropout_backward(g.data_ptr(), out.data_ptr(),
                                           None if bias is None else bias.data_ptr(), g.shape[0],
                                           g.shape[1], dropout_p, seed,
                                           None if seed_tensor is None else seed_tensor.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
    _capi.check(status, 'msda_dropout_backward')
    return out, bias


def _as_f32_2d(t, width):
    t = t.reshape(-1, width)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class Linear256Function(Function):
    """y = residual + dropout(x W^T + b) with the forward, the input-gradient GEMM,
    the weight gradient and the bias gradient on hand-written kernels: the three
    GEMMs on the tcgen05 tensor cores (3xTF32 split, fp32 accumulation in tensor
    memory), the padding mask, the storage dtype, dropout and the residual folded
    into the epilogue.

    mask_mode 1: masked rows of y are zero (mask after the projection,
    multi_scale_deform_attn.py:369-371); 2: the input rows are treated as zero,
    so y = bias there (mask before it, transformer.py:1706-1711).
    residual / dropout_p: `identity + dropout(output_proj(x))`,
    multi_scale_deform_attn.py:406-412 (not combinable with a mask)."""

    @staticmethod
    def forward(ctx, x, weight, bias, row_mask, mask_mode, out_dtype, residual=None, dropout_p=0.0):
        shape = x.shape
        n_out, n_in = weight.shape
        x2d = x.reshape(-1, n_in).contiguous()
        weight = weight.contiguous()
        mask_u8 = None
        if row_mask is not None and mask_mode:
            mask_u8 = row_mask.reshape(-1).to(torch.uint8).contiguous()
            if mask_u8.numel() != x2d.shape[0]:
                raise RuntimeError('row_mask has %d entries for %d rows' % (mask_u8.numel(), x2d.shape[0]))
        if mask_u8 is not None and (residual is not None or dropout_p > 0):
            raise RuntimeError('linear256: row_mask cannot be combined with residual / dropout')
        res2d = None
        if residual is not None:
            if residual.shape != shape[:-1] + (n_out,) or out_dtype != torch.float32:
                raise RuntimeError('linear256: residual must have the output shape; fp32 output only')
            res2d = _as_f32_2d(residual, n_out)
        seed, seed_t = _draw_seed() if dropout_p > 0 else (0, None)
        y = _linear_fused_raw(x2d, weight, None if bias is None else bias.contiguous(), mask_u8,
                              mask_mode if mask_u8 is not None else 0, dropout_p=dropout_p, seed=seed,
                              residual=res2d, out_dtype=out_dtype, seed_tensor=seed_t)
        ctx.save_for_backward(x2d, weight, mask_u8)
        ctx.mask_mode = mask_mode if mask_u8 is not None else 0
        ctx.has_bias = bias is not None
        ctx.x_shape = shape
        ctx.dropout = (dropout_p, seed, seed_t)
        ctx.has_residual = residual is not None
        return y.view(*shape[:-1], n_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x2d, weight, mask_u8 = ctx.saved_tensors
        n_out, n_in = weight.shape
        g = _as_f32_2d(grad_y, n_out)
        grad_x = grad_w = grad_b = grad_res = None
        if ctx.has_residual and ctx.needs_input_grad[6]:
            grad_res = grad_y
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        dropout_p, seed, seed_t = ctx.dropout
        if dropout_p > 0:
            # undo the epilogue's dropout (same keep decisions); the pass also sums the bias gradient
            g, grad_b = _dropout_backward_raw(g, dropout_p, seed, want_b, seed_t)
        elif want_b:
            # column sums of dY in one streaming pass (masked rows skipped in mode 1)
            grad_b = _colsum_raw(g, mask_u8 if ctx.mask_mode == 1 else None)
        if ctx.needs_input_grad[0]:
            # dX = dY W: the same kernel with W^T as the weight; masked rows get no gradient
            grad_x = _linear_fused_raw(g, weight.t().contiguous(), None, mask_u8,
                                       1 if ctx.mask_mode else 0).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            # dW = dY^T X, split-K on the tensor cores; mode 1 drops the masked rows of dY
            # (their outputs were forced to zero), mode 2 those of X (their inputs were)
            grad_w = _wgrad_raw(g, x2d, mask_u8, ctx.mask_mode, n_out, n_in)
        return grad_x, grad_w, grad_b, None, None, None, grad_res, None


def linear256(x, weight, bias=None, row_mask=None, mask_mode=0, out_dtype=torch.float32,
              residual=None, dropout_p=0.0):
    """Functional front-end of `Linear256Function` (see there)."""
    return Linear256Function.apply(x, weight, bias, row_mask, mask_mode, out_dtype, residual, dropout_p)


class FusedFFNFunction(Function):
    """identity + dropout(fc2(dropout(relu(fc1(x))))) -- the feed-forward block of the
    transformer layers (mmcv/cnn/bricks/transformer.py:1110-1120, 2 fcs, ReLU) -- as two
    tensor-core GEMMs forward (bias + ReLU + dropout, and bias + dropout + residual, in
    the epilogues) and four backward (the ReLU / dropout backward and the residual
    gradient in the epilogues too), plus two streaming bias-gradient passes."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, dropout_p, identity, add_identity):
        # identity None with add_identity: the residual is x itself
        shape = x.shape
        n_hidden, n_in = w1.shape
        x2d = _as_f32_2d(x, n_in)
        w1, w2 = w1.contiguous(), w2.contiguous()
        res2d = None
        if add_identity:
            res2d = x2d if identity is None else _as_f32_2d(identity, w2.shape[0])
        seed1, seed_t = _draw_seed() if dropout_p > 0 else (0, None)
        seed2, _ = _draw_seed() if dropout_p > 0 else (0, None)
        h = _linear_fused_raw(x2d, w1, None if b1 is None else b1.contiguous(), relu=True,
                              dropout_p=dropout_p, seed=seed1, seed_tensor=seed_t)
        y = _linear_fused_raw(h, w2, None if b2 is None else b2.contiguous(), dropout_p=dropout_p,
                              seed=seed2, residual=res2d, seed_tensor=seed_t)
        ctx.save_for_backward(x2d, h, w1, w2)
        ctx.seed_tensor = seed_t
        ctx.cfg = (shape, dropout_p, seed2, b1 is not None, b2 is not None,
                   add_identity and identity is not None, add_identity and identity is None)
        return y.view(*shape[:-1], w2.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        x2d, h, w1, w2 = ctx.saved_tensors
        shape, p, seed2, has_b1, has_b2, has_identity, identity_is_x = ctx.cfg
        n_hidden, n_in = w1.shape
        n_out = w2.shape[0]
        g = _as_f32_2d(grad_y, n_out)
        # through dropout 2 (+ bias gradient of fc2 from the same pass)
        if p > 0:
            g2, grad_b2 = _dropout_backward_raw(g, p, seed2, has_b2, ctx.seed_tensor)
        else:
            g2, grad_b2 = g, (_colsum_raw(g) if has_b2 else None)
        grad_w2 = _wgrad_raw(g2, h, None, 0, n_out, n_hidden)
        # dZ = (dY2 W2) * [h > 0] / (1 - p): h > 0 iff the ReLU passed and dropout 1 kept
        dz = _linear_fused_raw(g2, w2.t().contiguous(), gate=h, gate_scale=1.0 / (1.0 - p))
        grad_b1 = _colsum_raw(dz) if has_b1 else None
        grad_w1 = _wgrad_raw(dz, x2d, None, 0, n_hidden, n_in)
        grad_x = grad_identity = None
        if ctx.needs_input_grad[0]:
            # dX = dZ W1 (+ the identity branch's gradient when the identity is x itself)
            grad_x = _linear_fused_raw(dz, w1.t().contiguous(),
                                       residual=g if identity_is_x else None).view(shape)
        if has_identity and not identity_is_x:
            grad_identity = grad_y
        return grad_x, grad_w1, grad_b1, grad_w2, grad_b2, None, grad_identity, None


def ffn_supported(x, w1, w2):
    """True when `fused_ffn` has kernels for these tensors (fp32 CUDA, widths 128 / 256 / 1024)."""
    return (linear256_supported(x, w1) and w2.is_cuda and w2.dtype == torch.float32
            and w2.dim() == 2 and w2.shape[1] == w1.shape[0] and _linear_shape_ok(*w2.shape))


def fused_ffn(x, w1, b1, w2, b2, dropout_p=0.0, identity=None, add_identity=True):
    """identity + dropout(fc2(dropout(relu(fc1(x))))); identity=None means x itself;
    add_identity=False leaves the residual out."""
    return FusedFFNFunction.apply(x, w1, b1, w2, b2, float(dropout_p), identity, bool(add_identity))


# ---------------------------------------------------------------------------
# LayerNorm(256): the `norm` step after every attention module / feed-forward block
# ---------------------------------------------------------------------------
def layer_norm_supported(x, weight, bias):
    """True when `layer_norm256` has kernels: fp32 CUDA, 256 channels, affine."""
    return (x.is_cuda and x.dtype == torch.float32 and x.shape[-1] == 256 and x.numel() > 0
            and weight is not None and bias is not None and weight.dtype == torch.float32
            and weight.numel() == 256 and bias.numel() == 256)


class LayerNorm256Function(Function):
    """nn.LayerNorm(256) forward / backward as two streaming kernels (layernorm.cu): the
    backward writes grad_x and both parameter gradients in one pass over (x, grad_y)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        lib = _capi.load()
        x2d = _as_f32_2d(x, 256)
        weight, bias = weight.contiguous(), bias.contiguous()
        rows = x2d.shape[0]
        with torch.cuda.device(x2d.device):
            y = torch.empty_like(x2d)
            stats = torch.empty((2, rows), dtype=torch.float32, device=x2d.device)
            status = lib.msda_layernorm_forward(
                x2d.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), stats[0].data_ptr(),
                stats[1].data_ptr(), rows, 256, eps, torch.cuda.current_stream().cuda_stream)
        _capi.check(status, 'msda_layernorm_forward')
        ctx.save_for_backward(x2d, weight, stats)
        ctx.x_shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y):
        lib = _capi.load()
        x2d, weight, stats = ctx.saved_tensors
        g = _as_f32_2d(grad_y, 256)
        with torch.cuda.device(g.device):
            grad_x = torch.empty_like(x2d)
            grad_wb = torch.empty((2, 256), dtype=torch.float32, device=g.device)
            status = lib.msda_layernorm_backward(
                x2d.data_ptr(), g.data_ptr(), weight.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                grad_x.data_ptr(), grad_wb[0].data_ptr(), grad_wb[1].data_ptr(), x2d.shape[0], 256,
                torch.cuda.current_stream().cuda_stream)
        _capi.check(status, 'msda_layernorm_backward')
        return grad_x.view(ctx.x_shape), grad_wb[0], grad_wb[1], None


def layer_norm256(x, weight, bias, eps=1e-5):
    """Functional front-end of `LayerNorm256Function`."""
    return LayerNorm256Function.apply(x, weight, bias, float(eps))
