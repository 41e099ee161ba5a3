"""Build recipe for the C-ABI CUDA library (`libpavenet_msda.so`).

Compiles `pavenet_b200/csrc/*.cu` for sm_100a with nvcc into
`pavenet_b200/lib/` — in-tree, so the built library travels with the
repository snapshot.  nvcc cross-compiles without a GPU, so this also runs on
a CPU-only build host.

    python -m pavenet_b200._build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_DIR = os.path.join(_HERE, 'lib')
# PAVENET_MSDA_LIB selects an alternative build of the same ABI (kernel-tuning experiments)
LIB_PATH = os.environ.get('PAVENET_MSDA_LIB') or os.path.join(LIB_DIR, 'libpavenet_msda.so')
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), 'include')

SOURCES = ['msda_fwd.cu', 'msda_fwd_tile.cu', 'msda_bwd.cu', 'msda_bwd_priv.cu', 'msda_flat.cu', 'linear256_tc.cu', 'layernorm.cu', 'msda_capi.cu']
HEADERS = ['msda_common.cuh', 'msda_bwd_io.cuh', 'msda_kernels.h']

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libpavenet_msda.so')
    return exe


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(INCLUDE_DIR, 'pavenet_msda.h'))
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    """Compile the library if missing or older than its sources.

    Returns the path of the shared library.
    """
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    base = [_nvcc()] + NVCC_FLAGS + os.environ.get('PAVENET_MSDA_NVCC_EXTRA', '').split()
    if verbose:
        base += ['-Xptxas', '-v']
    objdir = os.path.join(LIB_DIR, 'obj%d' % os.getpid())
    os.makedirs(objdir, exist_ok=True)
    tmp = LIB_PATH + '.tmp%d' % os.getpid()
    try:
        # one nvcc per translation unit, side by side; then one link
        procs = []
        for src in SOURCES:
            obj = os.path.join(objdir, src[:-3] + '.o')
            cmd = base + ['-c', '-o', obj, os.path.join(CSRC, src)]
            procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                     stderr=subprocess.STDOUT, text=True)))
        log, failed = [], []
        for src, obj, proc in procs:
            out, _ = proc.communicate()
            log.append(out)
            if proc.returncode != 0:
                failed.append(src)
        if verbose or failed:
            sys.stderr.write(''.join(log))
        if failed:
            raise RuntimeError('nvcc failed on %s:\n%s' % (failed, ''.join(log)[-4000:]))
        link = [_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', tmp]
        link += [obj for _, obj, _ in procs]
        proc = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError('link failed (exit %d):\n%s' % (proc.returncode, proc.stdout[-4000:]))
        os.replace(tmp, LIB_PATH)
    finally:
        shutil.rmtree(objdir, ignore_errors=True)
        if os.path.exists(tmp):
            os.remove(tmp)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
