"""ctypes binding of the C ABI declared in `include/pavenet_msda.h`.

The shared library is the product; there is no Python or CPU fallback.  If it
is missing, importing the op fails loudly and tells the user how to build it.
"""
import ctypes
import os

from . import _build

# dtype codes of include/pavenet_msda.h
MSDA_F32, MSDA_F64, MSDA_BF16 = 0, 1, 2
MSDA_OK = 0

#: every symbol include/pavenet_msda.h declares (tests check the export list)
EXPORTED_SYMBOLS = (
    'msda_abi_version', 'msda_last_error', 'msda_launch_count', 'msda_launch_count_family',
    'msda_set_option',
    'msda_kernel_name', 'msda_forward', 'msda_forward_clear', 'msda_backward', 'msda_fused_forward',
    'msda_fused_backward', 'msda_linear256', 'msda_linear256_wgrad',
    'msda_colsum256', 'msda_linear_fused', 'msda_dropout_backward',
    'msda_layernorm_forward', 'msda_layernorm_backward',
    'msda_workspace_create', 'msda_workspace_destroy', 'msda_workspace_set_piece_bytes',
    'msda_host_alloc',
    'msda_host_free', 'msda_forward_host',
    'msda_forward_backward_host', 'msda_forward_backward_host_async', 'msda_workspace_wait',
)

_lib = None


class MsdaLibraryError(RuntimeError):
    """The CUDA library is missing, unloadable, or returned an error."""


def _declare(lib):
    c_int, c_void_p, c_i64p = ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p
    lib.msda_abi_version.restype = c_int
    lib.msda_abi_version.argtypes = []
    lib.msda_last_error.restype = ctypes.c_char_p
    lib.msda_last_error.argtypes = []
    lib.msda_launch_count.restype = ctypes.c_uint64
    lib.msda_launch_count.argtypes = []
    lib.msda_kernel_name.restype = ctypes.c_char_p
    lib.msda_kernel_name.argtypes = [c_int, c_int, c_int]
    lib.msda_forward.restype = c_int
    lib.msda_forward.argtypes = (
        [c_void_p, c_i64p, c_i64p, c_void_p, c_void_p, c_void_p] +
        [c_int] * 7 + [c_int, c_int, c_void_p])
    lib.msda_set_option.restype = c_int
    lib.msda_set_option.argtypes = [ctypes.c_char_p, c_int]
    lib.msda_launch_count_family.restype = ctypes.c_uint64
    lib.msda_launch_count_family.argtypes = [c_int]
    lib.msda_forward_clear.restype = c_int
    lib.msda_forward_clear.argtypes = (
        [c_void_p, c_i64p, c_i64p, c_void_p, c_void_p, c_void_p] +
        [c_int] * 7 + [c_int, c_int, c_void_p, ctypes.c_size_t, c_void_p])
    lib.msda_backward.restype = c_int
    lib.msda_backward.argtypes = (
        [c_void_p, c_i64p, c_i64p, c_void_p, c_void_p, c_void_p, c_void_p,
         c_void_p, c_void_p] + [c_int] * 7 + [c_int, c_int, c_int, c_void_p])
    lib.msda_fused_forward.restype = c_int
    lib.msda_fused_forward.argtypes = ([c_void_p] * 9 + [c_int] * 9 +
                                       [c_void_p, ctypes.c_size_t, c_void_p])
    lib.msda_fused_backward.restype = c_int
    lib.msda_fused_backward.argtypes = [c_void_p] * 14 + [c_int] * 9 + [c_void_p]
    lib.msda_linear256.restype = c_int
    lib.msda_linear256.argtypes = [c_void_p] * 4 + [c_int, c_void_p] + [c_int] * 4 + [c_void_p, c_void_p]
    lib.msda_linear256_wgrad.restype = c_int
    lib.msda_linear256_wgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.msda_linear_fused.restype = c_int
    lib.msda_linear_fused.argtypes = ([c_void_p] * 4 + [c_int, c_int, c_void_p, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_uint64, c_void_p, c_void_p, c_void_p] + [c_int] * 4 +
                                     [c_void_p, c_void_p])
    lib.msda_dropout_backward.restype = c_int
    lib.msda_dropout_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, ctypes.c_float,
                                          ctypes.c_uint64, c_void_p, c_void_p]
    lib.msda_layernorm_forward.restype = c_int
    lib.msda_layernorm_forward.argtypes = [c_void_p] * 6 + [c_int, c_int, ctypes.c_float, c_void_p]
    lib.msda_layernorm_backward.restype = c_int
    lib.msda_layernorm_backward.argtypes = [c_void_p] * 8 + [c_int, c_int, c_void_p]
    lib.msda_colsum256.restype = c_int
    lib.msda_colsum256.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]
    lib.msda_workspace_create.restype = c_int
    lib.msda_workspace_create.argtypes = [ctypes.POINTER(c_void_p)]
    lib.msda_workspace_destroy.restype = None
    lib.msda_workspace_destroy.argtypes = [c_void_p]
    lib.msda_workspace_set_piece_bytes.restype = c_int
    lib.msda_workspace_set_piece_bytes.argtypes = [c_void_p, ctypes.c_size_t]
    lib.msda_host_alloc.restype = c_void_p
    lib.msda_host_alloc.argtypes = [ctypes.c_size_t]
    lib.msda_host_free.restype = None
    lib.msda_host_free.argtypes = [c_void_p]
    lib.msda_forward_host.restype = c_int
    lib.msda_forward_host.argtypes = (
        [c_void_p] + [c_void_p] * 6 + [c_int] * 7 + [c_int, c_int])
    lib.msda_forward_backward_host.restype = c_int
    lib.msda_forward_backward_host.argtypes = (
        [c_void_p] + [c_void_p] * 10 + [c_int] * 7 + [c_int, c_int])
    lib.msda_forward_backward_host_async.restype = c_int
    lib.msda_forward_backward_host_async.argtypes = lib.msda_forward_backward_host.argtypes
    lib.msda_workspace_wait.restype = c_int
    lib.msda_workspace_wait.argtypes = [c_void_p]


def load():
    """Load (once) and return the ctypes handle of libpavenet_msda.so."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        raise MsdaLibraryError(
            'pavenet_b200: %s is missing. Build it with '
            '`python -m pavenet_b200._build` (needs nvcc); there is no CPU or '
            'PyTorch fallback for this op.' % path)
    try:
        lib = ctypes.CDLL(path)
    except OSError as exc:
        raise MsdaLibraryError('pavenet_b200: cannot load %s: %s' % (path, exc))
    missing = [s for s in EXPORTED_SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise MsdaLibraryError('pavenet_b200: %s lacks symbols %s (stale build?)'
                               % (path, missing))
    _declare(lib)
    _lib = lib
    return lib


def check(status, what):
    """Raise RuntimeError (as the reference's TORCH_CHECK does) on failure."""
    if status != MSDA_OK:
        msg = load().msda_last_error().decode('utf-8', 'replace')
        raise RuntimeError('%s failed (status %d): %s' % (what, status, msg))


def launch_count():
    return int(load().msda_launch_count())


#: kernel families of msda_launch_count_family (include/pavenet_msda.h)
KERNEL_FAMILIES = ('fwd_generic', 'bwd_generic', 'fwd_rows', 'bwd_rows', 'fwd_rows_fused',
                   'bwd_rows_fused', 'fwd_flat', 'bwd_flat', 'fwd_flat_fused', 'bwd_flat_fused',
                   'linear', 'linear_wgrad', 'colsum', 'layernorm', 'fwd_tile')


def family_counts():
    """{family name: launches since load} -- lets a test assert which kernel ran."""
    lib = load()
    return {name: int(lib.msda_launch_count_family(i)) for i, name in enumerate(KERNEL_FAMILIES)}


def set_option(name, value):
    """Kernel-selection knob (include/pavenet_msda.h, msda_set_option)."""
    check(load().msda_set_option(name.encode(), int(value)), 'msda_set_option')


def kernel_name(channels, dtype, value_dtype):
    return load().msda_kernel_name(channels, dtype, value_dtype).decode()
