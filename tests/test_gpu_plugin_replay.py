"""Plugin mode meets the real library on a GPU: the C-ABI call streams the UNMODIFIED reference produced while
driving the b200 backend (recorded in the build container on the CPU stand-in device, tests/abi_trace.py,
tools/gen_plugin_traces.py) are replayed call by call against libslf_b200.so, and every device -> host copy
-- macroscopic fields at output steps, packed ghost-layer buffers of the Collect* kernels (continuous and
sparse) that the reference ships to its neighbours -- is compared with what the oracle-backed device returned.
Single subdomains (BGK/MRT, AA/AB, PBC, pressure nodes, binary Shan-Chen) and all eight processes of a
2 x 2 x 2 decomposition (regtest/subdomains/3d_ldc.py layout)."""
import glob
import os

import numpy as np
import pytest

import abi_trace
import util

pytestmark = pytest.mark.gpu

TRACES = sorted(glob.glob(os.path.join(util.GOLD, 'trace_*.pkl.gz')))


@pytest.mark.parametrize('path', TRACES, ids=[os.path.basename(p)[6:-7] for p in TRACES])
def test_reference_call_stream_on_gpu_matches_oracle(path):
    from sailfish_b200 import backend_b200 as bk
    lib = bk.lib()
    lib.call('slf_init', 0)
    ops = abi_trace.load_fixture(path)
    res = abi_trace.replay(ops, lib)
    assert res, 'trace has no device -> host copy to compare'
    worst = 0.0
    for n, name, err, mask_ok, nfin in res:
        assert mask_ok, (n, name, 'finite masks differ')
        # max |gpu - oracle| / max(largest value of the copy, 1e-2): see abi_trace.Replayer.step
        assert err <= 2e-5, (n, name, err)
        worst = max(worst, err)
    print('%s: %d copies compared, worst relative deviation %.2e' % (os.path.basename(path), len(res), worst))
