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


def _zoo(name):
    import geometry_zoo
    from test_host_geometry import _zoo_namespace
    return geometry_zoo.make_cases(_zoo_namespace())[name]


ZOO_SPLITS = [
    # (zoo set-up, access pattern, split, global size): walls / outlets cut by, or next to, the subdomain faces.
    # Not here: link-bounce obstacles (half-way bounce-back, TMS) INSIDE the fluid that a subdomain face cuts
    # through.  The face transfer copies whole planes whatever the node types, so the stale populations of the
    # neighbour's wall-interior nodes overwrite the slots a wall node next to the face bounced into -- in the
    # reference as well (its connection code knows geometry, not node types; its decomposition regtests use full-way
    # walls).  Such set-ups depend on the decomposition by construction; walls on the domain boundary do not.
    ('slip_box3d', 'AA', ['--blocks=2,2,2'], (11, 10, 9)),
    ('slip_box3d_q27', 'AB', ['--blocks=1,2,2'], (11, 10, 9)),
    ('copy_duct3d', 'AB', ['--blocks=2,2,1'], (12, 9, 8)),
    ('yu_duct3d_mrt', 'AB', ['--blocks=3,1,2'], (12, 9, 8)),
    ('copy_channel2d', 'AB', ['--blocks=3,2'], (18, 11)),
    ('zouhe_box3d', 'AA', ['--blocks=2,2,2'], (12, 10, 9)),
    ('pressure_box3d', 'AB', ['--blocks=2,1,2'], (12, 10, 9)),
]


@pytest.mark.parametrize('name,ap,split,gsize', ZOO_SPLITS, ids=['%s-%s' % (z[0], z[1]) for z in ZOO_SPLITS])
def test_zoo_blocks_equal_single(fake, name, ap, split, gsize):
    """tests/geometry_zoo.py set-ups (TMS, slip, copy / Yu outflow, Zou-He, half-way bounce-back, pressure planes) on
    block decompositions whose faces cut through the walls and run next to the outlets: orientation and link-tag
    detection across subdomain faces, per-subdomain type remapping and unused tag bits, and the inward-neighbour reads
    of the outflow nodes must give the single-subdomain fields bit for bit."""
    cls, zargs = _zoo(name)
    geo = BlockSubdomainsGeometry2D if len(gsize) == 2 else BlockSubdomainsGeometry3D
    args = zargs + ['--access_pattern=' + ap, '--visc=0.06']
    _, one = block_util.run_blocks(cls, geo, args + ['--blocks=' + ','.join(['1'] * len(gsize))], 8)
    _, many = block_util.run_blocks(cls, geo, args + split, 8)
    np.testing.assert_array_equal(block_util.merge(one, gsize), block_util.merge(many, gsize))
    assert len(many) > 1


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('name,gsize,n', [('tms_channel2d', (16, 14), 2),      # (3 parts would cut the block on the wall)
                                          ('slip_channel2d', (16, 11), 2), ('slip_channel2d', (16, 11), 3)])
def test_2d_channel_periodic_in_x_split_in_y(fake, name, gsize, n, ap):
    """2-D, x periodic inside every subdomain, neighbours across the y faces (plane push): populations that leave
    through a y face AND the periodic x face must be wrapped inside the outgoing ghost row before it is pushed.  The
    reference's 2-D periodic kernel stops at the last real row (kernel_utils.mako:143-147; its own exchange carries
    those populations through the connection geometry instead); the linked variant of the kernel extends the rule
    on the linked side, as the 3-D rule does by itself.  Walls on the domain boundary (TMS: link bounce; slip) run
    along the periodic axis and across the split."""
    cls, zargs = _zoo(name)
    args = zargs + ['--access_pattern=' + ap, '--visc=0.06']
    _, one = block_util.run_blocks(cls, EqualSubdomainsGeometry2D, args + ['--subdomains=1'], 9)
    _, many = block_util.run_blocks(cls, EqualSubdomainsGeometry2D, args + ['--subdomains=%d' % n], 9)
    np.testing.assert_array_equal(block_util.merge(one, gsize), block_util.merge(many, gsize))
    assert len(many) == n and 'slf_halo_push' in fake.calls and 'slf_apply_periodic_boundary_conditions' in fake.calls


def _periodic_box3d(per, walls, grid):
    from sailfish_b200 import node_type as nt
    from sailfish_b200.lb_single import LBFluidSim
    from sailfish_b200.subdomain import Subdomain3D

    class Sub(Subdomain3D):
        def boundary_conditions(self, hx, hy, hz):
            w = np.zeros(np.broadcast(hx, hy, hz).shape, bool)
            if 'x' in walls:
                w |= (hx == 0) | (hx == self.gx - 1)
            if 'y' in walls:
                w |= (hy == 0) | (hy == self.gy - 1)
            if 'z' in walls:
                w |= (hz == 0) | (hz == self.gz - 1)
            if w.any():
                self.set_node(w, nt.NTFullBBWall)

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 1e-3 * np.sin(2 * np.pi * hx / self.gx) * np.cos(2 * np.pi * hz / self.gz)
            sim.vx[:] = 0.05 * np.sin(2 * np.pi * hy / self.gy)
            sim.vy[:] = 0.03 * np.cos(2 * np.pi * hz / self.gz)
            sim.vz[:] = 0.02 * np.sin(2 * np.pi * (hx + hy) / self.gx)

    class Sim(LBFluidSim):
        subdomain = Sub

        @classmethod
        def update_defaults(cls, d):
            d.update(lat_nx=10, lat_ny=9, lat_nz=12, grid=grid, **dict(('periodic_' + a, True) for a in per))
    return Sim


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('grid', ['D3Q19', 'D3Q27'])
@pytest.mark.parametrize('per,walls', [('xyz', ''), ('xy', 'z'), ('x', 'yz'), ('y', 'xz'), ('xz', 'y')])
def test_locally_periodic_axes_inside_a_z_split(fake, per, walls, grid, ap):
    """Every combination of locally periodic x / y (periodic kernels) with neighbours across z (plane push; a periodic z
    is closed by a wrap-around link), D3Q19 and D3Q27.  D3Q27 in the AA pattern is the case that needs
    pbc_landing_zone: the reference's edge entries for (±1, ±1, ±1) wrap along z whether or not z is periodic, and
    after in-place steps that stray copy would land in the ghost layer the neighbour delivers into."""
    sim = _periodic_box3d(per, walls, grid)
    args = ['--access_pattern=' + ap, '--visc=0.05']
    _, one = block_util.run_blocks(sim, EqualSubdomainsGeometry3D, args + ['--subdomains=1'], 8)
    _, many = block_util.run_blocks(sim, EqualSubdomainsGeometry3D, args + ['--subdomains=3'], 8)
    np.testing.assert_array_equal(block_util.merge(one, (10, 9, 12)), block_util.merge(many, (10, 9, 12)))
    assert 'slf_halo_push' in fake.calls


def _open_duct3d(wall_name):
    from sailfish_b200 import node_type as nt
    from sailfish_b200.lb_single import LBFluidSim
    from sailfish_b200.subdomain import Subdomain3D
    wall = getattr(nt, wall_name)

    class Sub(Subdomain3D):
        """Walls one node thick on the y and z faces of the domain, equilibrium-velocity inlet, do-nothing outlet."""
        def boundary_conditions(self, hx, hy, hz):
            w = (hy == 0) | (hy == self.gy - 1) | (hz == 0) | (hz == self.gz - 1)
            self.set_node(w, wall)
            free = np.logical_not(w)
            self.set_node((hx == 0) & free, nt.NTEquilibriumVelocity((0.03, 0.0, 0.0)))
            self.set_node((hx == self.gx - 1) & free, nt.NTDoNothing)

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0
            sim.vx[:] = 0.03

    class Sim(LBFluidSim):
        subdomain = Sub

        @classmethod
        def update_defaults(cls, d):
            d.update(lat_nx=12, lat_ny=9, lat_nz=10, grid='D3Q19')
    return Sim


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('wall', ['NTFullBBWall', 'NTHalfBBWall', 'NTWallTMS'])
@pytest.mark.parametrize('geo,one,split', [('equal', ['--subdomains=1'], ['--subdomains=3']),
                                           ('blocks', ['--blocks=1,1,1'], ['--blocks=2,2,2'])])
def test_open_duct_with_boundary_sensitive_nodes(fake, geo, one, split, wall, ap):
    """Do-nothing outlet and link-bounce walls ON THE DOMAIN BOUNDARY, cut by the subdomain faces.  These nodes keep
    what they find in the slots no node streams into, so a face push must not carry the never-written ghost entries
    next to the domain boundary across (slf_plan_set_face_links; the reference's partial connection maps): such
    subdomains exchange through the windowed face push even when they are only split along z."""
    sim = _open_duct3d(wall)
    g = EqualSubdomainsGeometry3D if geo == 'equal' else BlockSubdomainsGeometry3D
    args = ['--access_pattern=' + ap, '--visc=0.06']
    _, a = block_util.run_blocks(sim, g, args + one, 10)
    ctrl, b = block_util.run_blocks(sim, g, args + split, 10)
    np.testing.assert_array_equal(block_util.merge(a, (12, 9, 10)), block_util.merge(b, (12, 9, 10)))
    assert all(r._general for r in ctrl.runners) and 'slf_plan_set_face_links' in fake.calls


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('geo,one,split', [('equal', ['--subdomains=1'], ['--subdomains=3']),
                                           ('equal', ['--subdomains=1'], ['--subdomains=2', '--conn_axis=x']),
                                           ('blocks', ['--blocks=1,1'], ['--blocks=3,2'])])
def test_channel_with_do_nothing_outlet_on_splits(fake, geo, one, split, ap):
    cls, zargs = _zoo('channel2d')
    g = EqualSubdomainsGeometry2D if geo == 'equal' else BlockSubdomainsGeometry2D
    args = zargs + ['--access_pattern=' + ap, '--visc=0.06']
    _, a = block_util.run_blocks(cls, g, args + one, 10)
    _, b = block_util.run_blocks(cls, g, args + split, 10)
    np.testing.assert_array_equal(block_util.merge(a, (20, 12)), block_util.merge(b, (20, 12)))
