"""Host logic of the self-contained runner (sailfish_b200/{controller,subdomain_runner,runtime,io}.py) on
the CPU: the device behind the C ABI is tests/fake_device.py (host memory + the oracle), so what is exercised
is everything above it -- slab scheduling and plane pushes between subdomains, the link-aware PBC ordering,
the Shan-Chen phases, step replay, checkpoint / restore and the two output paths.  (The same scenarios run on
a B200 in tests/test_gpu_*.py.)"""
import glob
import sys

import numpy as np
import pytest

import fake_device
from sailfish_b200.examples import ldc_2d, ldc_3d, poiseuille_3d, sc_separation_3d
from sailfish_b200.geo import EqualSubdomainsGeometry2D, EqualSubdomainsGeometry3D


@pytest.fixture
def fake(oracle, monkeypatch):
    return fake_device.install_fake(monkeypatch)


def run(cls, geo, args, steps):
    from sailfish_b200.controller import LBSimulationController
    ctrl = LBSimulationController(cls, geo)
    ctrl.setup(args + ['--quiet'])
    for r in ctrl.runners:
        sim, spec = r._sim, r._spec
        if hasattr(sim, 'phi'):
            # the example draws noise from the global generator: make the initial state a function of the global
            # coordinates instead, so that it does not depend on the decomposition
            hz, hy, hx = np.meshgrid(np.arange(spec.oz, spec.oz + spec.nz), np.arange(spec.ny), np.arange(spec.nx),
                                     indexing='ij')
            sim.rho[:] = 1.0 + 1e-3 * np.sin(0.7 * hx + 1.3 * hy + 0.4 * hz)
            sim.phi[:] = 1.0 + 1e-3 * np.cos(0.5 * hx - 0.9 * hy + 1.1 * hz)
            for b in (r._dev.d_rho, r._dev.d_phi):
                r._dev.be.to_buf(b)
            r._dev.set_initial_conditions()
    for k in range(steps):
        ctrl.step_all(k == steps - 1)
    ctrl.sync()
    out = []
    for r in sorted(ctrl.runners, key=lambda r: r._spec.id):
        r._fields_to_host()
        sim, sl = r._sim, r._spec._nonghost_slice
        names = ['rho'] + (['phi'] if hasattr(sim, 'phi') else [])
        out.append([np.array(r.field_base(getattr(sim, n)))[sl] for n in names] +
                   [np.array(r.field_base(c))[sl] for c in sim.v])
    return [np.concatenate([o[k] for o in out], axis=0) for k in range(len(out[0]))]


SPLIT = {
    'ldc3d': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D, ['--lat_nx=12', '--lat_ny=10', '--lat_nz=15', '--visc=0.04'], 3),
    'ldc2d': (ldc_2d.LDCSim, EqualSubdomainsGeometry2D, ['--lat_nx=16', '--lat_ny=14', '--visc=0.05'], 2),
    'pipe_force_x': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                     ['--lat_nx=12', '--lat_ny=10', '--lat_nz=12', '--flow_direction=x', '--drive=force'], 2),
    'sc3d': (sc_separation_3d.SeparationSCSim, EqualSubdomainsGeometry3D,
             ['--lat_nx=10', '--lat_ny=8', '--lat_nz=12', '--seed=3'], 3),
}


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('name', sorted(SPLIT))
def test_split_equals_single(fake, name, ap):
    cls, geo, args, n = SPLIT[name]
    args = args + ['--access_pattern=' + ap]
    np.random.seed(0)
    one = run(cls, geo, args + ['--subdomains=1'], 9)
    many = run(cls, geo, args + ['--subdomains=%d' % n], 9)
    for a, b in zip(one, many):
        np.testing.assert_array_equal(a, b)
    assert 'slf_halo_push' in fake.calls and 'slf_plan_set_links' in fake.calls


def test_step_replay_equals_stepping(fake):
    from sailfish_b200.controller import LBSimulationController
    res = []
    for replay in (True, False):
        ctrl = LBSimulationController(ldc_3d.LDCSim)
        ctrl.setup(['--lat_nx=10', '--lat_ny=9', '--lat_nz=8', '--visc=0.04', '--access_pattern=AA', '--quiet'])
        r = ctrl.runners[0]
        if replay:
            r.run_steps(7, save_last=True)
        else:
            for k in range(7):
                r.step(k == 6)
        r._fields_to_host()
        res.append(np.array(r.field_base(r._sim.rho)))
    np.testing.assert_array_equal(res[0], res[1])
    assert 'slf_run_steps' in fake.calls


CASES = {
    'ldc3d_AA': (ldc_3d.LDCSim, None, ['--lat_nx=12', '--lat_ny=10', '--lat_nz=9', '--visc=0.04', '--access_pattern=AA']),
    'ldc3d_AB_split': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D,
                       ['--lat_nx=12', '--lat_ny=10', '--lat_nz=12', '--visc=0.04', '--access_pattern=AB', '--subdomains=2']),
    'sc3d_AA': (sc_separation_3d.SeparationSCSim, EqualSubdomainsGeometry3D,
                ['--lat_nx=10', '--lat_ny=8', '--lat_nz=8', '--access_pattern=AA', '--seed=7']),
}


def launch(monkeypatch, cls, geo, args):
    from sailfish_b200.controller import LBSimulationController
    monkeypatch.setattr(sys, 'argv', ['x'] + args + ['--quiet'])
    return LBSimulationController(cls, geo).run()


@pytest.mark.parametrize('name', sorted(CASES))
def test_restart_invariance(fake, tmp_path, monkeypatch, name):
    """tests/gpu/checkpoint.sh of the reference, on the host logic: 7 steps + checkpoint + 6 more == 13 steps."""
    cls, geo, args = CASES[name]
    cp, one, two = str(tmp_path / 'cp'), str(tmp_path / 'one'), str(tmp_path / 'two')
    launch(monkeypatch, cls, geo, args + ['--max_iters=7', '--final_checkpoint', '--checkpoint_file=' + cp])
    assert glob.glob(cp + '.7.*.cpoint.npz')
    launch(monkeypatch, cls, geo, args + ['--max_iters=13', '--restore_from=' + cp + '.7', '--output=' + two, '--every=13',
                                          '--sync_output'])
    launch(monkeypatch, cls, geo, args + ['--max_iters=13', '--output=' + one, '--every=13', '--sync_output'])
    f1, f2 = sorted(glob.glob(one + '.*.13.npz')), sorted(glob.glob(two + '.*.13.npz'))
    assert f1 and len(f1) == len(f2)
    for a, b in zip(f1, f2):
        ra, rb = np.load(a), np.load(b)
        for k in ra.files:
            np.testing.assert_array_equal(ra[k], rb[k])


@pytest.mark.parametrize('name', sorted(CASES))
def test_fast_output_equals_reference_style_output(fake, tmp_path, monkeypatch, name):
    cls, geo, args = CASES[name]
    files = {}
    for tag, extra in (('fast', []), ('sync', ['--sync_output'])):
        out = str(tmp_path / tag)
        launch(monkeypatch, cls, geo, args + ['--max_iters=8', '--every=2', '--output=' + out] + extra)
        files[tag] = sorted(glob.glob(out + '.*.npz'))
    assert files['fast'] and [f.split('/')[-1][4:] for f in files['fast']] == [f.split('/')[-1][4:] for f in files['sync']]
    for a, b in zip(files['fast'], files['sync']):
        fa, fb = np.load(a), np.load(b)
        assert sorted(fa.files) == sorted(fb.files)
        for k in fa.files:
            np.testing.assert_array_equal(fa[k], fb[k])
    assert 'slf_pack_field' in fake.calls


def test_random_decompositions(fake):
    """Seeded sample of random sizes / models / lattices / subdomain counts (uneven remainders, thin slabs):
    split == single, bit for bit."""
    import random
    rng = random.Random(20240924)
    for _ in range(16):
        kind = rng.choice(['ldc3d', 'ldc2d', 'pipe', 'sc3d'])
        ap = rng.choice(['AA', 'AB'])
        steps = rng.randint(1, 6)
        if kind == 'ldc3d':
            nz = rng.randint(6, 18)
            args = ['--lat_nx=%d' % rng.randint(4, 12), '--lat_ny=%d' % rng.randint(4, 10), '--lat_nz=%d' % nz, '--visc=0.04']
            args += rng.choice([['--model=bgk'], ['--model=mrt'], ['--grid=D3Q27']])
            cls, geo, outer = ldc_3d.LDCSim, EqualSubdomainsGeometry3D, nz
        elif kind == 'ldc2d':
            outer = rng.randint(6, 22)
            args = ['--lat_nx=%d' % rng.randint(5, 36), '--lat_ny=%d' % outer, '--visc=0.05', '--model=' + rng.choice(['bgk', 'mrt'])]
            cls, geo = ldc_2d.LDCSim, EqualSubdomainsGeometry2D
        elif kind == 'pipe':
            outer = rng.randint(8, 16)
            args = ['--lat_nx=%d' % rng.randint(8, 12), '--lat_ny=%d' % rng.randint(8, 12), '--lat_nz=%d' % outer,
                    '--flow_direction=' + rng.choice('xyz'), '--drive=' + rng.choice(['force', 'pressure']),
                    '--model=' + rng.choice(['bgk', 'mrt'])]
            cls, geo = poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D
        else:
            outer = rng.randint(6, 12)
            args = ['--lat_nx=%d' % rng.randint(4, 9), '--lat_ny=%d' % rng.randint(4, 9), '--lat_nz=%d' % outer, '--seed=1',
                    '--G11=%g' % rng.choice([0.0, -0.2])]
            cls, geo = sc_separation_3d.SeparationSCSim, EqualSubdomainsGeometry3D
        nsub = rng.randint(2, max(2, min(5, outer // 2)))
        args += ['--access_pattern=' + ap]
        one = run(cls, geo, args + ['--subdomains=1'], steps)
        many = run(cls, geo, args + ['--subdomains=%d' % nsub], steps)
        for a, b in zip(one, many):
            np.testing.assert_array_equal(a, b, err_msg='%s %s nsub=%d steps=%d' % (kind, args, nsub, steps))
