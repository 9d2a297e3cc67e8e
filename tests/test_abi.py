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
