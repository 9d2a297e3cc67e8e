"""N > 1 path on CPU: two ranks (gloo), one subdomain each, planes exchanged point to point;
the merged fields must equal the single-subdomain run (decomposition invariance)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import multi_util
import util
from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.examples import ldc_3d, poiseuille_3d

CASES = {
    'ldc3d': (ldc_3d.LDCSim, ['--lat_nx=10', '--lat_ny=9', '--lat_nz=14', '--visc=0.04']),
    'poi3d_force_z': (poiseuille_3d.PoiseuilleSim,
                      ['--lat_nx=10', '--lat_ny=10', '--lat_nz=12', '--flow_direction=z', '--drive=force']),
}


@pytest.mark.parametrize('name,ap', [('ldc3d', 'AA'), ('ldc3d', 'AB'), ('poi3d_force_z', 'AA')])
def test_two_ranks_equal_one_subdomain(oracle, tmp_path, name, ap):
    out = str(tmp_path / 'merged.npy')
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
           '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(util.ROOT, 'tests', '_gloo_worker.py'), name, ap, out]
    env = dict(os.environ, OMP_NUM_THREADS='1')
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    merged = np.load(out)
    cls, args = CASES[name]
    _, one = multi_util.build_case(cls, EqualSubdomainsGeometry3D, args + ['--access_pattern=' + ap, '--subdomains=1'])
    s1 = multi_util.run_local(oracle, one, 11)
    wet = one[0]['sd'].fluid_map()
    ref = [multi_util.merge_field(one, s1, lambda s: s.o.rho, 3)] + \
          [multi_util.merge_field(one, s1, (lambda d: (lambda s: s.o.v[d]))(d), 3) for d in range(3)]
    for a, b in zip(ref, merged):
        assert np.allclose(a[wet], b[wet], rtol=0, atol=1e-13)


BLOCK_CASES = [('ldc3d', 'AA', '2,2,1', 4), ('ldc3d', 'AB', '2,2,1', 4), ('ldc3d', 'AA', '2,1,2', 4), ('ldc3d_q27', 'AA', '2,2,2', 8),
               ('ldc2d', 'AA', '2,2', 4)]


@pytest.mark.parametrize('name,ap,blocks,nranks', BLOCK_CASES)
def test_blocks_by_message_passing_equal_one_subdomain(oracle, tmp_path, name, ap, blocks, nranks):
    """Block decompositions with the dimension-ordered face exchange as point-to-point MESSAGES between ranks that
    share no memory (one rank per block): the merged fields equal the single-subdomain run.  Edge and corner data
    travel in two (D3Q27: three) hops."""
    sys.path.insert(0, os.path.join(util.ROOT, 'tests'))
    import _gloo_block_worker as w
    out = str(tmp_path / 'merged.npy')
    port = 29300 + (os.getpid() % 200) + 7 * len(blocks)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % nranks,
           '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(util.ROOT, 'tests', '_gloo_block_worker.py'), name, ap, blocks, out]
    env = dict(os.environ, OMP_NUM_THREADS='1')
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    merged = np.load(out)
    cls, geo, args = w.CASES[name]
    ones = ','.join(['1'] * len(blocks.split(',')))
    _, one = multi_util.build_case(cls, geo, args + ['--access_pattern=' + ap, '--blocks=' + ones])
    s1 = multi_util.run_local(oracle, one, 9)
    part = one[0]
    o = s1[part['spec'].id].o
    sl = part['spec']._nonghost_slice
    ref = np.array([o.rho.reshape(part['sd'].arr_shape)[sl]] + [c.reshape(part['sd'].arr_shape)[sl] for c in o.v])
    wet = part['sd'].fluid_map()
    assert merged.shape == ref.shape
    assert np.allclose(ref[:, wet], merged[:, wet], rtol=0, atol=1e-13)
