"""Builds libslf_b200.so in-tree with nvcc for sm_100a (B200).  ``python -m sailfish_b200.build``."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libslf_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-shared', '-Xptxas', '-v']


def sources():
    return [os.path.join(SRC, f) for f in sorted(os.listdir(SRC)) if f.endswith(('.cu', '.cuh'))] + \
           [os.path.join(os.path.dirname(HERE), 'include', 'slf_b200.h')]


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError('nvcc not found at %s and %s is missing or stale' % (NVCC, LIB))
    cmd = [NVCC] + FLAGS + ['-o', LIB, os.path.join(SRC, 'slf_api.cu')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + log)
    if res.returncode != 0:
        sys.stderr.write(log[-6000:])
        raise RuntimeError('nvcc failed (%d)' % res.returncode)
    if verbose:
        print(log)
    return LIB


if __name__ == '__main__':
    build(force='-f' in sys.argv, verbose='-v' in sys.argv)
    print(LIB)
