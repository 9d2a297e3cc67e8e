"""The reference's own multi-subdomain regression tests (regtest/subdomains/*.py, SURVEY.md section 8c
items 3 and 4), run UNCHANGED on the b200 backend through sailfish_b200.plugin:

* 2d_/3d_propagation.py -- literal known-answer vectors for streaming across subdomain faces, edges and
  corners, periodic and partially periodic boxes, misaligned blocks, AA with 2 and 4 subdomains;
* 2d_/3d_ldc.py (3 / 4 / 8 blocks), 2d_cylinder.py -- decomposition invariance to 6 decimals;
* 2d_binary.py, binary_pbc.py -- the Shan-Chen mixture on 2 and 4 blocks and in shifted periodic boxes
  (macro-field halo kernels); 3d_sphere.py -- flow past a sphere on several blocks;
* tests/gpu/do_nothing_node.py -- AA == AB with the do-nothing outflow node.

The whole reference stack runs as is (controller, machine master, one OS process per subdomain, ZeroMQ
connectors, Collect/Distribute{Continuous,Sparse}[Macro]Data[WithSwap] through the C entry points'
argument contracts); the device behind the C ABI is tests/fake_device.py here (no GPU in the build
container, no reference checkout on the GPU box).  Tests on lattices / models outside SURVEY section 8
(D3Q15, free energy) must fail with UnsupportedConfiguration, not silently."""
import json
import os
import subprocess
import sys

import pytest

import util

REF = os.environ.get('SAILFISH_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'regtest', 'subdomains')),
                                reason='reference tree not available')

RUNS = {
    '3d_propagation': ('AB', None),
    '2d_propagation': ('AB', None),
    '3d_ldc': ('AA', 100),
    '2d_ldc': ('AA', 100),
    '2d_cylinder': ('AB', 100),
    '2d_binary': ('AB', 100),
    '3d_sphere': ('AB', 100),
    'binary_pbc': ('AB', None),
    'tests/gpu/do_nothing_node.py': ('AB', None),     # AA == AB with the do-nothing outflow node, 1000 steps
}
if os.environ.get('SLF_ALL_REGTESTS'):
    RUNS.update({'3d_ldc_ab': ('AB', 100), '2d_ldc_ab': ('AB', 100)})
OUT_OF_SCOPE = ('Q15', 'FreeEnergy', 'FEInterblock')


@pytest.fixture(scope='module')
def runs(oracle):
    """All regtest modules run concurrently, each in its own interpreter."""
    procs = {}
    for key, (ap, iters) in RUNS.items():
        cmd = [sys.executable, os.path.join(util.ROOT, 'tests', 'reference_regtest_runner.py'), REF,
               key.replace('_ab', ''), ap] + ([str(iters)] if iters else [])
        procs[key] = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return procs


@pytest.mark.parametrize('key', sorted(RUNS))
def test_reference_regtest_passes_on_b200_backend(runs, key):
    out, err = runs[key].communicate(timeout=900)
    lines = [l for l in out.splitlines() if l.startswith('{')]
    assert lines, err[-3000:]
    outcomes = json.loads(lines[-1])
    assert outcomes, err[-3000:]
    in_scope = dict((k, v) for k, v in outcomes.items() if not any(t in k for t in OUT_OF_SCOPE))
    assert in_scope, outcomes
    bad = dict((k, v) for k, v in in_scope.items() if v != 'ok')
    assert not bad, (bad, err[-3000:])
    for k, v in outcomes.items():
        if 'FE' in k or 'FreeEnergy' in k:
            assert v != 'ok' and 'UnsupportedConfiguration' in err, (k, v)
