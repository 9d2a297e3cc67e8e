import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def _gpu_available():
    """A CUDA device and the built library (asked through the C ABI itself, not through torch)."""
    try:
        import ctypes
        from sailfish_b200 import backend_b200 as bk
        if os.path.exists('/dev/nvidiactl') and not os.path.exists(bk.LIB_PATH):
            return True, ''                      # a GPU box without the library must fail loudly, not skip
        if not os.path.exists(bk.LIB_PATH):
            return False, 'libslf_b200.so is not built and there is no CUDA device'
        n = ctypes.c_int(0)
        lib = ctypes.CDLL(bk.LIB_PATH)
        lib.slf_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
        if lib.slf_device_count(ctypes.byref(n)) != 0 or n.value == 0:
            return False, 'no CUDA device'
        return True, ''
    except Exception as e:                       # noqa: BLE001
        return False, repr(e)


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of erroring with B200Error."""
    if not any('gpu' in it.keywords for it in items):
        return
    ok, why = _gpu_available()
    if ok:
        return
    skip = pytest.mark.skip(reason='gpu test: ' + why)
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)
