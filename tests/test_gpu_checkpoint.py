"""Restart invariance (the reference's tests/gpu/checkpoint.sh): run 13 steps, write a checkpoint, restore it
and continue to 20 steps; the result must equal an uninterrupted 20-step run bit for bit.  13 is odd on purpose
(as the reference's 103): in the AA pattern the raw lattice layout after an odd number of steps differs from the
one after an even number, so this also pins the layout round trip (subdomain_runner.py:1414-1449)."""
import glob
import importlib
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    'ldc3d_AA': ('ldc_3d', 'LDCSim', None, ['--lat_nx=20', '--lat_ny=18', '--lat_nz=17', '--visc=0.04', '--access_pattern=AA']),
    'ldc3d_AB': ('ldc_3d', 'LDCSim', None, ['--lat_nx=20', '--lat_ny=18', '--lat_nz=17', '--visc=0.04', '--access_pattern=AB']),
    'ldc2d_AA': ('ldc_2d', 'LDCSim', None, ['--lat_nx=40', '--lat_ny=33', '--visc=0.05', '--access_pattern=AA']),
    'pipe_force_AA': ('poiseuille_3d', 'PoiseuilleSim', 'EqualSubdomainsGeometry3D',
                      ['--lat_nx=18', '--lat_ny=18', '--lat_nz=20', '--flow_direction=z', '--drive=force',
                       '--visc=0.1', '--access_pattern=AA']),
    'sc3d_AA': ('sc_separation_3d', 'SeparationSCSim', 'EqualSubdomainsGeometry3D',
                ['--lat_nx=16', '--lat_ny=12', '--lat_nz=10', '--access_pattern=AA', '--seed=7']),
    'sc3d_AB': ('sc_separation_3d', 'SeparationSCSim', 'EqualSubdomainsGeometry3D',
                ['--lat_nx=16', '--lat_ny=12', '--lat_nz=10', '--access_pattern=AB', '--seed=7']),
    'ldc3d_AA_split': ('ldc_3d', 'LDCSim', 'EqualSubdomainsGeometry3D',
                       ['--lat_nx=20', '--lat_ny=18', '--lat_nz=17', '--visc=0.04', '--access_pattern=AA',
                        '--subdomains=2']),
}


def run(monkeypatch, mod, clsname, geoname, args):
    from sailfish_b200 import geo as geomod
    from sailfish_b200.controller import LBSimulationController
    monkeypatch.setattr(sys, 'argv', ['x'] + args + ['--quiet'])
    return LBSimulationController(getattr(mod, clsname), getattr(geomod, geoname) if geoname else None).run()


@pytest.mark.parametrize('name', sorted(CASES))
def test_restart_invariance(tmp_path, monkeypatch, name):
    modname, clsname, geoname, args = CASES[name]
    mod = importlib.import_module('sailfish_b200.examples.' + modname)
    cp, one, two = str(tmp_path / 'cp'), str(tmp_path / 'one'), str(tmp_path / 'two')
    run(monkeypatch, mod, clsname, geoname, args + ['--max_iters=13', '--final_checkpoint', '--checkpoint_file=' + cp])
    files = glob.glob(cp + '.13.*.cpoint.npz')
    assert files
    saved = np.load(files[0])
    assert 'state' in saved.files and 'dist0a' in saved.files
    assert ('dist0b' in saved.files) == ('AB' in name)
    run(monkeypatch, mod, clsname, geoname, args + ['--max_iters=20', '--restore_from=' + cp + '.13', '--output=' + two,
                                                     '--every=20', '--sync_output'])
    run(monkeypatch, mod, clsname, geoname, args + ['--max_iters=20', '--output=' + one, '--every=20', '--sync_output'])
    f1, f2 = sorted(glob.glob(one + '.*.20.npz')), sorted(glob.glob(two + '.*.20.npz'))
    assert f1 and len(f1) == len(f2)
    for a, b in zip(f1, f2):
        ra, rb = np.load(a), np.load(b)
        for k in ra.files:
            np.testing.assert_array_equal(ra[k], rb[k])
