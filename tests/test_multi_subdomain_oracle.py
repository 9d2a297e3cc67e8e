"""Decomposition invariance -- the reference's own multi-subdomain bar
(regtest/subdomains/{2d,3d}_ldc.py: N subdomains == 1 subdomain to 6 decimals,
regtest/subdomains/util.py:27-30): the z/y-split stepping + plane exchange used by the B200
runner, executed with the CPU oracle, must reproduce the single-subdomain run."""
import numpy as np
import pytest

import multi_util
from sailfish_b200.geo import EqualSubdomainsGeometry2D, EqualSubdomainsGeometry3D
from sailfish_b200.examples import ldc_2d, ldc_3d, poiseuille_3d

CASES = {
    'ldc3d': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D,
              ['--lat_nx=10', '--lat_ny=9', '--lat_nz=14', '--visc=0.04'], 3),
    'ldc3d_mrt': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D,
                  ['--lat_nx=10', '--lat_ny=9', '--lat_nz=14', '--visc=0.04', '--model=mrt'], 2),
    'ldc2d': (ldc_2d.LDCSim, EqualSubdomainsGeometry2D, ['--lat_nx=12', '--lat_ny=11', '--visc=0.05'], 2),
    'poi3d_force_z': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                      ['--lat_nx=10', '--lat_ny=10', '--lat_nz=12', '--flow_direction=z', '--drive=force'], 2),
    'poi3d_force_x': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                      ['--lat_nx=12', '--lat_ny=10', '--lat_nz=12', '--flow_direction=x', '--drive=force'], 3),
    'poi3d_press_z': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                      ['--lat_nx=10', '--lat_ny=10', '--lat_nz=16', '--flow_direction=z', '--drive=pressure',
                       '--stationary', '--model=mrt'], 4),
}


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('ap', ['AB', 'AA'])
def test_n_subdomains_equal_one(oracle, name, ap):
    cls, geo, args, n = CASES[name]
    steps = 11
    args = args + ['--access_pattern=' + ap]
    _, one = multi_util.build_case(cls, geo, args + ['--subdomains=1'])
    _, many = multi_util.build_case(cls, geo, args + ['--subdomains=%d' % n])
    assert len(many) == n
    s1 = multi_util.run_local(oracle, one, steps)
    sn = multi_util.run_local(oracle, many, steps)
    dim = one[0]['spec'].dim
    wet1 = one[0]['sd'].fluid_map()
    for getter in [lambda s: s.o.rho] + [(lambda d: (lambda s: s.o.v[d]))(d) for d in range(dim)]:
        a = multi_util.merge_field(one, s1, getter, dim)
        b = multi_util.merge_field(many, sn, getter, dim)
        assert a.shape == b.shape
        assert np.allclose(a[wet1], b[wet1], rtol=0, atol=1e-13), np.abs(a[wet1] - b[wet1]).max()
