"""Drop-in check: the UNMODIFIED reference (controller -> master -> SubdomainRunner.run, its own
examples, its own output writer) runs on ``--backends=b200`` through sailfish_b200.plugin.

The build container has the reference but no GPU (and the GPU box has no reference), so the
device behind the C ABI is replaced here by tests/fake_device.py (host memory + the CPU oracle).
What this pins is everything between the reference and the C entry points: backend duck type,
kernel-name/argument mapping (lb_single.py:72-185), iteration patching (backend_cuda.py:128-130),
AA/AB lattice selection, PBC kernel variants and the "file N holds step N's pre-collision
moments" convention.  The .npz files the reference writes must match the golden vectors.
Skipped where /root/reference does not exist."""
import glob
import os
import sys

import numpy as np
import pytest

import util

REF = os.environ.get('SAILFISH_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'sailfish')),
                                reason='reference tree not available')

CASES = {
    'ldc3d_bgk': ('examples/ldc_3d.py', 'LDCSim', None,
                  ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04'], 20),
    'ldc3d_mrt': ('examples/ldc_3d.py', 'LDCSim', None,
                  ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04', '--model=mrt'], 20),
    'ldc2d_bgk': ('examples/ldc_2d.py', 'LDCSim', None, ['--lat_nx=16', '--lat_ny=14', '--visc=0.05'], 20),
    'poi3d_force_bgk': ('examples/poiseuille_3d.py', 'PoiseuilleSim', 'EqualSubdomainsGeometry3D',
                        ['--grid=D3Q19', '--lat_nx=14', '--lat_ny=12', '--lat_nz=12', '--flow_direction=x',
                         '--drive=force', '--visc=0.1'], 20),
    'poi3d_press_mrt': ('examples/poiseuille_3d.py', 'PoiseuilleSim', 'EqualSubdomainsGeometry3D',
                        ['--grid=D3Q19', '--lat_nx=12', '--lat_ny=12', '--lat_nz=16', '--flow_direction=z',
                         '--drive=pressure', '--stationary', '--model=mrt', '--visc=0.1'], 20),
    'box3d_bgk': None,   # needs a custom class, see below
}


@pytest.fixture(scope='module')
def reference():
    from sailfish_b200 import plugin
    plugin.install(REF)
    sys.path.insert(0, os.path.join(util.ROOT, 'tools'))
    import ref_harness
    return ref_harness


@pytest.mark.parametrize('name', [k for k, v in CASES.items() if v is not None])
@pytest.mark.parametrize('ap', ['AB', 'AA'])
def test_reference_runs_on_b200_backend(reference, oracle, monkeypatch, tmp_path, name, ap):
    import fake_device
    fake = fake_device.install_fake(monkeypatch)
    script, cls_name, geo_name, args, steps = CASES[name]
    mod = reference.load_example(script)
    from sailfish.controller import LBSimulationController
    from sailfish import geo
    out = str(tmp_path / 'out')
    argv = [script] + args + ['--access_pattern=' + ap, '--backends=b200', '--debug_single_process',
                              '--max_iters=%d' % steps, '--every=%d' % steps, '--output=' + out, '--quiet']
    monkeypatch.setattr(sys, 'argv', argv)
    ctrl = LBSimulationController(getattr(mod, cls_name), getattr(geo, geo_name) if geo_name else None)
    ctrl.run()
    assert 'slf_collide_and_propagate' in fake.calls and 'slf_set_initial_conditions' in fake.calls
    files = sorted(glob.glob(out + '.0.*%d.npz' % steps))
    assert files, glob.glob(out + '*')
    res = np.load(files[-1])
    c = util.load_case(name + '_' + ap)
    m = util.real_mask(c)
    wet = c['fluid_map']
    rho = c['rho'][m].reshape(res['rho'].shape)
    assert np.isnan(res['rho'][~wet]).all()          # io.py:53-59 masks non-fluid nodes
    err = np.abs(res['rho'][wet] - rho[wet]).max() / np.abs(rho[wet]).max()
    assert err < 5e-6, err
    vscale = max(np.abs(c['v'][d][m].reshape(rho.shape)[wet]).max() for d in range(res['v'].shape[0]))
    for d in range(res['v'].shape[0]):
        v = c['v'][d][m].reshape(rho.shape)
        assert np.abs(res['v'][d][wet] - v[wet]).max() / vscale < 5e-5
    if c['periodic'].any():
        assert 'slf_apply_periodic_boundary_conditions' in fake.calls


@pytest.mark.parametrize('ap', ['AB', 'AA'])
def test_reference_shan_chen_runs_on_b200_backend(reference, oracle, monkeypatch, tmp_path, ap):
    """BASELINE config 5: the reference's examples/binary_fluid/sc_separation_3d.py, unmodified, through
    ShanChenPrepareMacroFields / ShanChenCollideAndPropagate{0,1} / ApplyMacroPeriodicBoundaryConditions
    (lb_binary.py:41-128,423-517) on the b200 backend; output vs the golden vector of the same run."""
    import fake_device
    fake = fake_device.install_fake(monkeypatch)
    script = 'examples/binary_fluid/sc_separation_3d.py'
    mod = reference.load_example(script)
    from sailfish.controller import LBSimulationController
    from sailfish.geo import LBGeometry3D
    steps = 12
    out = str(tmp_path / 'out')
    argv = [script, '--lat_nx=12', '--lat_ny=10', '--lat_nz=9', '--seed=1234', '--access_pattern=' + ap,
            '--backends=b200', '--debug_single_process', '--max_iters=%d' % steps, '--every=%d' % steps,
            '--output=' + out, '--quiet']
    monkeypatch.setattr(sys, 'argv', argv)
    LBSimulationController(mod.SeparationSCSim, LBGeometry3D).run()
    for name in ('slf_sc_set_initial_conditions', 'slf_sc_prepare_macro_fields', 'slf_sc_collide_and_propagate',
                 'slf_apply_macro_periodic_boundary_conditions', 'slf_apply_periodic_boundary_conditions'):
        assert name in fake.calls, name
    files = sorted(glob.glob(out + '.0.*%d.npz' % steps))
    assert files, glob.glob(out + '*')
    res = np.load(files[-1])
    c = np.load(os.path.join(util.ROOT, 'tests', 'golden', 'steps_sc3d_%s.npz' % ap))
    inner = (slice(1, -1),) * 3
    for k in ('rho', 'phi'):
        want = c[k][inner][..., :12]
        assert res[k].shape == want.shape
        assert np.abs(res[k] - want).max() / np.abs(want).max() < 5e-6, k


NATIVE = {
    'ldc3d': ('examples/ldc_3d.py', 'LDCSim', 'ldc_3d', None, ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04']),
    'ldc3d_mrt': ('examples/ldc_3d.py', 'LDCSim', 'ldc_3d', None,
                  ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04', '--model=mrt']),
    'ldc2d': ('examples/ldc_2d.py', 'LDCSim', 'ldc_2d', None, ['--lat_nx=16', '--lat_ny=14', '--visc=0.05']),
    'pipe_pressure': ('examples/poiseuille_3d.py', 'PoiseuilleSim', 'poiseuille_3d', 'EqualSubdomainsGeometry3D',
                      ['--grid=D3Q19', '--lat_nx=12', '--lat_ny=12', '--lat_nz=16', '--flow_direction=z', '--drive=pressure',
                       '--stationary', '--model=mrt', '--visc=0.1']),
    'pipe_force': ('examples/poiseuille_3d.py', 'PoiseuilleSim', 'poiseuille_3d', 'EqualSubdomainsGeometry3D',
                   ['--grid=D3Q19', '--lat_nx=14', '--lat_ny=12', '--lat_nz=12', '--flow_direction=x', '--drive=force',
                    '--visc=0.1']),
    'sc3d': ('examples/binary_fluid/sc_separation_3d.py', 'SeparationSCSim', 'sc_separation_3d', 'LBGeometry3D',
             ['--lat_nx=12', '--lat_ny=10', '--lat_nz=9', '--seed=1234']),
}


@pytest.mark.parametrize('name', sorted(NATIVE))
@pytest.mark.parametrize('ap', ['AB', 'AA'])
def test_native_host_stack_equals_reference_host_stack(reference, oracle, monkeypatch, tmp_path, name, ap):
    """The same set-up through the two host stacks above the C ABI -- the unmodified reference (its example,
    controller, runner, output writer) and this package's own (sailfish_b200.examples, controller, runner) --
    on the same (fake) device: the output files must be identical, bit for bit, NaN mask included."""
    import importlib
    import fake_device
    fake_device.install_fake(monkeypatch)
    script, cls_name, native_mod, geo_name, args = NATIVE[name]
    steps = 10
    # both stacks hand a context dict under the reference's keys to params.from_context; keep what comes out
    from sailfish_b200 import params as P
    made = []
    real_from_context = P.from_context

    def recording_from_context(ctx):
        p = real_from_context(ctx)
        vals = []
        for fname, ftype in p._fields_:
            v = getattr(p, fname)
            if fname == 'node_params':
                v = [v[i] for i in range(p.num_node_params)]
            elif hasattr(v, '__len__'):
                v = list(v)
            vals.append((fname, v))
        made.append((vals, set(ctx)))
        return p
    monkeypatch.setattr(P, 'from_context', recording_from_context)
    common = args + ['--access_pattern=' + ap, '--max_iters=%d' % steps, '--every=%d' % steps, '--quiet']
    # reference stack
    mod = reference.load_example(script)
    from sailfish.controller import LBSimulationController as RefController
    from sailfish import geo as refgeo
    out_ref = str(tmp_path / 'ref')
    monkeypatch.setattr(sys, 'argv', [script] + common + ['--backends=b200', '--debug_single_process', '--output=' + out_ref])
    RefController(getattr(mod, cls_name), getattr(refgeo, geo_name) if geo_name else None).run()
    n_ref = len(made)
    assert n_ref >= 1
    # native stack
    from sailfish_b200 import geo as mygeo
    from sailfish_b200.controller import LBSimulationController
    nat = importlib.import_module('sailfish_b200.examples.' + native_mod)
    out_nat = str(tmp_path / 'nat')
    native_args = [a for a in common if not a.startswith('--grid=')] + ['--output=' + out_nat, '--sync_output']
    monkeypatch.setattr(sys, 'argv', ['x'] + native_args)
    geo_cls = getattr(mygeo, geo_name, None) if geo_name else None
    LBSimulationController(getattr(nat, cls_name), geo_cls).run()
    assert len(made) == n_ref + 1
    # the kernel parameter block is the same, field for field, and the mirror's context uses the reference's keys
    assert made[-1][0] == made[0][0]
    assert made[-1][1] <= made[0][1], sorted(made[-1][1] - made[0][1])
    fr = sorted(glob.glob(out_ref + '.0.*%d.npz' % steps))
    fn = sorted(glob.glob(out_nat + '.0.*%d.npz' % steps))
    assert fr and fn
    a, b = np.load(fr[-1]), np.load(fn[-1])
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        np.testing.assert_array_equal(a[k], b[k])


def _zoo_reference_namespace():
    from sailfish import node_type as rnt
    from sailfish.subdomain import Subdomain2D, Subdomain3D
    from sailfish.lb_single import LBFluidSim
    from sailfish.lb_base import LBForcedSim
    ns = dict(Subdomain2D=Subdomain2D, Subdomain3D=Subdomain3D, LBFluidSim=LBFluidSim, LBForcedSim=LBForcedSim)
    for n in dir(rnt):
        if n.startswith('NT'):
            ns[n] = getattr(rnt, n)
    return ns


@pytest.mark.parametrize('case', util.ZOO_STEP_CASES)
def test_reference_runs_geometry_zoo_on_b200_backend(reference, oracle, monkeypatch, tmp_path, case):
    """tests/geometry_zoo.py written against the REFERENCE's classes (its Subdomain, node types, LBFluidSim,
    controller, runner) on --backends=b200: every node class the plan accepts -- half-way bounce-back, TMS, slip,
    equilibrium / regularised / Zou-He planes, do-nothing, copy and Yu outflow -- reaches the C entry points through the
    reference's own context (from_context) and reproduces the step goldens."""
    import fake_device
    import geometry_zoo
    fake = fake_device.install_fake(monkeypatch)
    name, ap = case[len('zoo_'):-3], case[-2:]
    cls, args = geometry_zoo.make_cases(_zoo_reference_namespace())[name]
    d = {}
    cls.update_defaults(d)
    from sailfish.controller import LBSimulationController
    from sailfish.geo import LBGeometry2D, LBGeometry3D
    c = util.load_case(case)
    steps = int(c['steps'])
    out = str(tmp_path / 'out')
    argv = ['zoo'] + args + ['--access_pattern=' + ap, '--visc=0.06', '--backends=b200', '--debug_single_process',
                            '--max_iters=%d' % steps, '--every=%d' % steps, '--output=' + out, '--quiet']
    monkeypatch.setattr(sys, 'argv', argv)
    LBSimulationController(cls, LBGeometry2D if d['grid'] == 'D2Q9' else LBGeometry3D).run()
    assert 'slf_collide_and_propagate' in fake.calls
    files = sorted(glob.glob(out + '.0.*%d.npz' % steps))
    assert files, glob.glob(out + '*')
    res = np.load(files[-1])
    m = util.real_mask(c)
    wet = c['fluid_map']
    rho = c['rho'][m].reshape(res['rho'].shape)
    assert np.isnan(res['rho'][~wet]).all()
    err = np.abs(res['rho'][wet] - rho[wet]).max() / np.abs(rho[wet]).max()
    assert err < 5e-6, err
    vscale = max(np.abs(c['v'][k][m].reshape(rho.shape)[wet]).max() for k in range(res['v'].shape[0]))
    for k in range(res['v'].shape[0]):
        v = c['v'][k][m].reshape(rho.shape)
        assert np.abs(res['v'][k][wet] - v[wet]).max() / vscale < 1e-4
