"""Build the C-ABI CUDA library ``libtorchoptics_b200.so`` in-tree for sm_100a.

    python -m torchoptics_b200.build [--force]

One ``nvcc`` invocation, no torch headers: the library's interface is the plain
C ABI of ``include/torchoptics_b200.h``.  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libtorchoptics_b200.so')
SOURCES = [os.path.join(CSRC, 'trace_kernels.cu')]
HEADERS = [os.path.join(CSRC, 'trace_core.cuh'), os.path.join(CSRC, 'trace_core_asph.cuh'),
           os.path.join(CSRC, 'trace_kernels_gen.cuh'), os.path.join(CSRC, 'peer_exchange.cuh'),
           os.path.join(CSRC, 'spot_rev.cuh'), os.path.join(CSRC, 'psf_kernels.cuh'), os.path.join(CSRC, 'paraxial.cuh'),
           os.path.join(PKG, '..', 'include', 'torchoptics_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--fmad=true', '-Xcompiler', '-fPIC', '-shared', '-cudart', 'shared']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *(['-Xptxas', '-v'] if verbose else []), *SOURCES, '-o', LIB]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
