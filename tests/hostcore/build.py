"""Build the CPU test aid tests/hostcore/_hostcore.so (see hostcore.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_hostcore.so')


def build(force=False):
    src = os.path.join(HERE, 'hostcore.cpp')
    csrc = os.path.join(HERE, '..', '..', 'torchoptics_b200', 'csrc')
    deps = [src, os.path.join(csrc, 'trace_core.cuh'), os.path.join(csrc, 'trace_core_asph.cuh'),
            os.path.join(csrc, 'paraxial.cuh'), os.path.join(HERE, '..', '..', 'include', 'torchoptics_b200.h')]
    if (not force and os.path.exists(SO)
            and os.path.getmtime(SO) >= max(os.path.getmtime(d) for d in deps)):
        return SO
    subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-fno-fast-math', '-std=c++17',
                           '-shared', '-fPIC', '-x', 'c++', src, '-o', SO])
    return SO


if __name__ == '__main__':
    print(build(force=True))
