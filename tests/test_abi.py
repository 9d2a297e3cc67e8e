"""The C-ABI library loads and exports every symbol include/slf_b200.h declares (no compute
calls: this runs without a GPU), the ctypes struct matches the C struct, and the product
fails loudly -- never falls back to a CPU path -- when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

import util
from sailfish_b200 import backend_b200 as bk
from sailfish_b200 import params as P

HEADER = os.path.join(util.ROOT, 'include', 'slf_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(slf_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def library():
    from sailfish_b200 import build
    build.build()
    return ctypes.CDLL(bk.LIB_PATH)


def test_header_and_binding_agree():
    assert declared_symbols() == bk.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol(library):
    for name in declared_symbols():
        assert hasattr(library, name), name


def test_abi_version_and_struct_size(library):
    assert library.slf_abi_version() == 6
    # sizeof(SlfParams) as the C compiler sees it
    prog = '#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu", sizeof(SlfParams));return 0;}' % HEADER
    exe = os.path.join(util.ROOT, 'tests', '.sizeof_params')
    subprocess.run(['/usr/bin/gcc', '-x', 'c', '-', '-o', exe], input=prog.encode(), check=True)
    size = int(subprocess.check_output([exe]))
    os.unlink(exe)
    assert size == ctypes.sizeof(P.SlfParams)


def test_no_cpu_fallback():
    """Without a CUDA device the backend must raise, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(bk.B200Error):
        bk.B200Backend(None, 0)


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import, call, link or execute
    anything under oracle/."""
    import re
    pkg = os.path.join(util.ROOT, 'sailfish_b200')
    pat = re.compile(r'(^\s*(from|import)\s+oracle\b)|liboracle|oracle/|orc_[a-z_]+\s*\(', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), '%s references the oracle' % f


def test_face_windows_of_the_library_equal_the_host_reference(library):
    """slf_halo_face_window (the pure host function slf_halo_push_face takes its windows from) against
    tests/halo_ref.face_windows, which the CPU stand-in device uses: every lattice, face, step kind and set of
    linked faces.  No device needed."""
    import ctypes
    import itertools
    import halo_ref
    from sailfish_b200 import lattice
    fn = library.slf_halo_face_window
    fn.restype = ctypes.c_int
    I3, I2 = ctypes.c_int * 3, ctypes.c_int * 2
    for lid, name in enumerate(('D2Q9', 'D3Q19', 'D3Q27')):
        grid = lattice.get(name)
        D = grid.dim
        lat = [7, 6, 5][:D]
        for face, swap, links in itertools.product(range(2 * D), (0, 1), range(0, 64, 1 if D == 2 else 3)):
            want = dict((slot, (src, dst)) for slot, src, dst in
                        halo_ref.face_windows(grid, lat, face, bool(swap), lat, links))
            got = {}
            for i in range(grid.Q):
                lo, cnt = I2(0, 0), I2(0, 0)
                if fn(lid, I3(*(lat + [1] * (3 - D))), face, swap, links, i, lo, cnt):
                    got[int(grid.opposite[i]) if swap else i] = (list(lo), list(cnt))
            assert sorted(got) == sorted(want), (name, face, swap, links)
            axis = face // 2
            others = [a for a in range(D) if a != axis]
            for slot, (src, dst) in want.items():
                lo, cnt = got[slot]
                for b, ax in enumerate(others):
                    assert src[ax] == dst[ax] == slice(lo[b], lo[b] + cnt[b]), (name, face, swap, links, slot, ax, src, lo, cnt)
