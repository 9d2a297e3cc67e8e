"""Builds libslf_b200.so in-tree with nvcc for sm_100a (B200).  ``python -m sailfish_b200.build``.

One object per translation unit, compiled in parallel, then linked with nvcc -shared."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', 'obj')
LIB = os.path.join(HERE, 'libslf_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def units():
    return sorted(f for f in os.listdir(SRC) if f.endswith('.cu'))


def headers():
    return [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith('.cuh')] + \
           [os.path.join(os.path.dirname(HERE), 'include', 'slf_b200.h')]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def up_to_date():
    deps = headers() + [os.path.join(SRC, u) for u in units()]
    return not _stale(LIB, deps)


def _compile(unit):
    src = os.path.join(SRC, unit)
    obj = os.path.join(OBJ, unit[:-3] + '.o')
    if not _stale(obj, headers() + [src]):
        return unit, 0, '(cached) ' + unit
    cmd = [NVCC] + FLAGS + ['-c', '-o', obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return unit, res.returncode, ' '.join(cmd) + '\n' + res.stdout + res.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError('nvcc not found at %s and %s is missing or stale' % (NVCC, LIB))
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.unlink(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(_compile, units()))
    log = '\n'.join(r[2] for r in results)
    bad = [r for r in results if r[1] != 0]
    if not bad:
        objs = [os.path.join(OBJ, u[:-3] + '.o') for u in units()]
        cmd = [NVCC, '-shared', '-o', LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += '\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr
        if res.returncode != 0:
            bad = [('link', res.returncode, '')]
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write(log)
    if bad:
        sys.stderr.write(log[-8000:])
        raise RuntimeError('nvcc failed: %s' % ', '.join(b[0] for b in bad))
    if verbose:
        print(log)
    return LIB


if __name__ == '__main__':
    build(force='-f' in sys.argv, verbose='-v' in sys.argv)
    print(LIB)
