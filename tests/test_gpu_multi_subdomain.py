"""Several subdomains on one GPU, driven by one process through the public controller: the
peer-to-peer plane pushes (slf_halo_push) and the slab scheduling of SubdomainRunner.step must
reproduce the single-subdomain GPU run (the reference's regtest/subdomains bar: 6 decimals) and
agree with the oracle."""
import numpy as np
import pytest

import multi_util
from sailfish_b200.geo import EqualSubdomainsGeometry2D, EqualSubdomainsGeometry3D
from sailfish_b200.examples import ldc_2d, ldc_3d, poiseuille_3d

pytestmark = pytest.mark.gpu

CASES = {
    'ldc3d': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D, ['--lat_nx=20', '--lat_ny=18', '--lat_nz=30', '--visc=0.04'], 3),
    'ldc3d_mrt': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D,
                  ['--lat_nx=20', '--lat_ny=18', '--lat_nz=24', '--visc=0.04', '--model=mrt'], 2),
    'ldc2d': (ldc_2d.LDCSim, EqualSubdomainsGeometry2D, ['--lat_nx=24', '--lat_ny=22', '--visc=0.05'], 2),
    'poi3d_force_z': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                      ['--lat_nx=14', '--lat_ny=14', '--lat_nz=16', '--flow_direction=z', '--drive=force'], 2),
    'poi3d_force_x': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                      ['--lat_nx=16', '--lat_ny=12', '--lat_nz=18', '--flow_direction=x', '--drive=force'], 3),
    'poi3d_press_z': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                      ['--lat_nx=12', '--lat_ny=12', '--lat_nz=24', '--flow_direction=z', '--drive=pressure',
                       '--stationary', '--model=mrt'], 4),
}


def gpu_run(cls, geo, args, steps):
    from sailfish_b200.controller import LBSimulationController
    ctrl = LBSimulationController(cls, geo)
    ctrl.setup(args + ['--quiet'])
    for k in range(steps):
        ctrl.step_all(k == steps - 1)
    ctrl.sync()
    dim = ctrl.dim
    pieces = []
    for r in ctrl.runners:
        r._fields_to_host()
        sl = r._spec._nonghost_slice
        f = [np.array(r.field_base(r._sim.rho))[sl]] + [np.array(r.field_base(c))[sl] for c in r._sim.v]
        pieces.append((r._spec.location[dim - 1], f))
    pieces.sort(key=lambda t: t[0])
    return [np.concatenate([p[1][k] for p in pieces], axis=0) for k in range(1 + dim)]


@pytest.mark.parametrize('name', sorted(CASES))
@pytest.mark.parametrize('ap', ['AB', 'AA'])
def test_gpu_n_subdomains_equal_one(oracle, name, ap):
    cls, geo, args, n = CASES[name]
    steps = 11
    args = args + ['--access_pattern=' + ap]
    one = gpu_run(cls, geo, args + ['--subdomains=1'], steps)
    many = gpu_run(cls, geo, args + ['--subdomains=%d' % n], steps)
    _, parts = multi_util.build_case(cls, geo, args + ['--subdomains=1'])
    wet = parts[0]['sd'].fluid_map()
    ref = multi_util.run_local(oracle, parts, steps, dtype=np.float64)
    dim = parts[0]['spec'].dim
    refs = [multi_util.merge_field(parts, ref, lambda s: s.o.rho, dim)] + \
           [multi_util.merge_field(parts, ref, (lambda d: (lambda s: s.o.v[d]))(d), dim) for d in range(dim)]
    ref32 = multi_util.run_local(oracle, parts, steps, dtype=np.float32)
    refs32 = [multi_util.merge_field(parts, ref32, lambda s: s.o.rho, dim)] + \
             [multi_util.merge_field(parts, ref32, (lambda d: (lambda s: s.o.v[d]))(d), dim) for d in range(dim)]
    vscale = max(1e-30, max(np.abs(c[wet]).max() for c in refs[1:]))
    for k, (a, b, c, c32) in enumerate(zip(one, many, refs, refs32)):
        # identical kernels on identical data: the split run must match the single run bit for bit
        assert np.array_equal(a[wet], b[wet]), np.abs(a[wet] - b[wet]).max()
        # and stay inside the fp32 round-off envelope of the reference arithmetic (fp32 oracle vs fp64)
        scale = np.abs(c[wet]).max() if k == 0 else vscale      # ||.||_inf of rho, resp. of the velocity field
        err = np.abs(b[wet] - c[wet]).max() / scale
        env = np.abs(c32[wet] - c[wet]).max() / scale
        assert err <= 2.0 * env + 1e-6, (k, err, env)


@pytest.mark.parametrize('ap', ['AB', 'AA'])
@pytest.mark.parametrize('name,n', [('tms_channel2d', 2), ('slip_channel2d', 3)])
def test_gpu_2d_channel_periodic_in_x_split_in_y(name, n, ap):
    """2-D, x periodic inside every subdomain, neighbours across the y faces: the linked variant of the periodic
    kernel wraps the outgoing ghost row before it is pushed (tests/test_block_decomposition_fake.py has the same
    set-ups on the CPU stand-in device).  Split == single, bit for bit; the single run itself is covered by
    tests/test_gpu_zoo_node_types.py."""
    import geometry_zoo
    from test_host_geometry import _zoo_namespace
    cls, zargs = geometry_zoo.make_cases(_zoo_namespace())[name]
    args = zargs + ['--access_pattern=' + ap, '--visc=0.06']
    one = gpu_run(cls, EqualSubdomainsGeometry2D, args + ['--subdomains=1'], 9)
    many = gpu_run(cls, EqualSubdomainsGeometry2D, args + ['--subdomains=%d' % n], 9)
    for a, b in zip(one, many):
        np.testing.assert_array_equal(a, b)


@pytest.mark.xfail(strict=False, reason='pbc_landing_zone was mirrored into apply_pbc from the CPU stand-in after the '
                   "round's GPU minutes were spent: D3Q27 + AA + locally periodic axes inside a z-split has not run on "
                   'hardware yet (tests/test_block_decomposition_fake.py covers it on the CPU)')
@pytest.mark.parametrize('per,walls', [('xy', 'z'), ('x', 'yz')])
def test_gpu_d3q27_aa_locally_periodic_inside_z_split(per, walls):
    from test_block_decomposition_fake import _periodic_box3d
    sim = _periodic_box3d(per, walls, 'D3Q27')
    args = ['--access_pattern=AA', '--visc=0.05']
    one = gpu_run(sim, EqualSubdomainsGeometry3D, args + ['--subdomains=1'], 8)
    many = gpu_run(sim, EqualSubdomainsGeometry3D, args + ['--subdomains=3'], 8)
    for a, b in zip(one, many):
        np.testing.assert_array_equal(a, b)


@pytest.mark.xfail(strict=False, reason='windowed exchange for subdomains with do-nothing / link-bounce nodes '
                   '(slf_plan_set_face_links): host-side change made after the GPU minutes of the round were spent; the '
                   'window arithmetic of the library is checked on the CPU (tests/test_abi.py), the scheduling on the '
                   'CPU stand-in (tests/test_block_decomposition_fake.py); not yet run on hardware')
@pytest.mark.parametrize('ap', ['AB', 'AA'])
@pytest.mark.parametrize('wall', ['NTHalfBBWall', 'NTWallTMS'])
def test_gpu_open_duct_with_boundary_sensitive_nodes_split_in_z(wall, ap):
    from test_block_decomposition_fake import _open_duct3d
    sim = _open_duct3d(wall)
    args = ['--access_pattern=' + ap, '--visc=0.06']
    one = gpu_run(sim, EqualSubdomainsGeometry3D, args + ['--subdomains=1'], 10)
    many = gpu_run(sim, EqualSubdomainsGeometry3D, args + ['--subdomains=3'], 10)
    for a, b in zip(one, many):
        np.testing.assert_array_equal(a, b)
