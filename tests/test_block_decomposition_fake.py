"""Subdomains split along x and / or y as well (regtest/subdomains/{2d,3d}_ldc.py layouts) through the native
runner's dimension-ordered face exchange, on the CPU stand-in device: the merged fields must equal the
single-subdomain run bit for bit (the reference's bar is 6 decimals, regtest/subdomains/util.py:27-30).
Edge and corner data travel in two / three hops; D3Q27 exercises the corner hop."""
import numpy as np
import pytest

import block_util
import fake_device
from sailfish_b200.examples import ldc_2d, ldc_3d
from sailfish_b200.geo import (BlockSubdomainsGeometry2D, BlockSubdomainsGeometry3D, EqualSubdomainsGeometry2D,
                               EqualSubdomainsGeometry3D)


@pytest.fixture
def fake(oracle, monkeypatch):
    return fake_device.install_fake(monkeypatch)


L3 = ['--lat_nx=13', '--lat_ny=11', '--lat_nz=10', '--visc=0.04']
L2 = ['--lat_nx=17', '--lat_ny=14', '--visc=0.05']
CASES = {
    'ldc3d_2x2x2': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3, ['--blocks=2,2,2'], (13, 11, 10)),
    'ldc3d_3x1x2': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3, ['--blocks=3,1,2'], (13, 11, 10)),
    'ldc3d_1x2x1_mrt': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3 + ['--model=mrt'], ['--blocks=1,2,1'], (13, 11, 10)),
    'ldc3d_q27_2x2x2': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3 + ['--grid=D3Q27'], ['--blocks=2,2,2'], (13, 11, 10)),
    'ldc3d_conn_x': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D, L3, ['--subdomains=3', '--conn_axis=x'], (13, 11, 10)),
    'ldc3d_conn_y': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D, L3, ['--subdomains=2', '--conn_axis=y'], (13, 11, 10)),
    'ldc2d_2x2': (ldc_2d.LDCSim, BlockSubdomainsGeometry2D, L2, ['--blocks=2,2'], (17, 14)),
    'ldc2d_3x2_mrt': (ldc_2d.LDCSim, BlockSubdomainsGeometry2D, L2 + ['--model=mrt'], ['--blocks=3,2'], (17, 14)),
    'ldc2d_conn_x': (ldc_2d.LDCSim, EqualSubdomainsGeometry2D, L2, ['--subdomains=2', '--conn_axis=x'], (17, 14)),
}


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_blocks_equal_single(fake, name, ap):
    cls, geo, args, split, gsize = CASES[name]
    args = args + ['--access_pattern=' + ap]
    single_split = ['--blocks=' + ','.join(['1'] * len(gsize))] if any(a.startswith('--blocks') for a in split) else ['--subdomains=1']
    _, one = block_util.run_blocks(cls, geo, args + single_split, 9)
    _, many = block_util.run_blocks(cls, geo, args + split, 9)
    a, b = block_util.merge(one, gsize), block_util.merge(many, gsize)
    np.testing.assert_array_equal(a, b)
    assert len(many) > 1 and 'slf_halo_push_face' in fake.calls


def test_periodic_axis_split_wraps_around(fake):
    """A periodic axis that is split is closed by a wrap-around link instead of the PBC kernels."""
    from sailfish_b200.examples import poiseuille_3d
    args = ['--lat_nx=12', '--lat_ny=10', '--lat_nz=10', '--flow_direction=x', '--drive=force', '--access_pattern=AA']
    _, one = block_util.run_blocks(poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D, args + ['--subdomains=1'], 7)
    _, many = block_util.run_blocks(poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                                    args + ['--subdomains=3', '--conn_axis=x'], 7)
    np.testing.assert_array_equal(block_util.merge(one, (12, 10, 10)), block_util.merge(many, (12, 10, 10)))
