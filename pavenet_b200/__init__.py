"""pavenet_b200 — B200-native (sm_100a) multi-scale deformable attention for
PAVE-Net: the sampling op (forward + backward) behind a C ABI, and the host-side
mirror of the reference's operator / module interface around it.

Layout
  csrc/          hand-written CUDA kernels + the extern "C" boundary
                 (declared in ../include/pavenet_msda.h)
  lib/           the built shared library (git-ignored, built in-tree)
  _build.py      nvcc recipe           _capi.py   ctypes binding
  functional.py  MultiScaleDeformableAttnFunction / ext_module (reference names)
  modules.py     the attention module classes (reference names, fused multi-frame)
  registry.py    type-string registry glue / `install()` into a live mmcv

There is no CPU path: ops raise if the CUDA library is missing or a tensor is
not on a CUDA device.
"""
from . import _capi
from .functional import (LayerNorm256Function, layer_norm256, FusedFFNFunction, FusedMultiScaleDeformableAttnFunction, HostWorkspace,
                         Linear256Function, MultiScaleDeformableAttnFunction, ext_module,
                         ffn_supported, fused_ffn, fused_supported, linear256, linear256_supported,
                         fuse_frames_as_levels, ms_deform_attn_backward,
                         ms_deform_attn_forward)
from .modules import (FFN, LayerNorm, MulFramesMultiScaleDeformableAttentionNumFrames3,
                      MulFramesMultiScaleDeformableAttentionNumFrames5,
                      MulFramesMultiScaleDeformablePoseAttentionNumFrames3,
                      MulFramesMultiScaleDeformablePoseAttentionNumFrames5,
                      MultiScaleDeformableAttention,
                      MultiScaleDeformablePoseAttention)
from .registry import (ATTENTION, FEEDFORWARD_NETWORK, OPERA_ATTENTION, build_attention,
                       build_feedforward_network, install)

__version__ = '0.1.0'

__all__ = [
    'MultiScaleDeformableAttnFunction', 'ext_module', 'ms_deform_attn_forward',
    'ms_deform_attn_backward', 'fuse_frames_as_levels', 'HostWorkspace',
    'FusedMultiScaleDeformableAttnFunction', 'fused_supported', 'Linear256Function',
    'linear256', 'linear256_supported', 'FusedFFNFunction', 'fused_ffn', 'ffn_supported', 'FFN',
    'FEEDFORWARD_NETWORK', 'build_feedforward_network', 'LayerNorm', 'LayerNorm256Function', 'layer_norm256',
    'MultiScaleDeformableAttention', 'MultiScaleDeformablePoseAttention',
    'MulFramesMultiScaleDeformablePoseAttentionNumFrames3',
    'MulFramesMultiScaleDeformablePoseAttentionNumFrames5',
    'MulFramesMultiScaleDeformableAttentionNumFrames3',
    'MulFramesMultiScaleDeformableAttentionNumFrames5',
    'ATTENTION', 'OPERA_ATTENTION', 'build_attention', 'install',
]
