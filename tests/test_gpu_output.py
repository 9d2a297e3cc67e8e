"""Output fast path (slf_pack_field + asynchronous writer, sailfish_b200/io.py) against the
reference-style path (synchronous copy of the padded fields, host-side NaN mask of
LBOutput.mask_nonfluid_nodes built from the host geometry, savez in the main thread)."""
import glob

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


CASES = {
    'ldc3d': ('ldc_3d', 'LDCSim', None, ['--lat_nx=20', '--lat_ny=18', '--lat_nz=17', '--visc=0.04',
                                          '--access_pattern=AA', '--max_iters=20', '--every=4']),
    'ldc2d': ('ldc_2d', 'LDCSim', None, ['--lat_nx=40', '--lat_ny=33', '--visc=0.05', '--access_pattern=AB',
                                          '--max_iters=12', '--every=3']),
    'pipe': ('poiseuille_3d', 'PoiseuilleSim', 'EqualSubdomainsGeometry3D',
             ['--lat_nx=18', '--lat_ny=18', '--lat_nz=20', '--flow_direction=z', '--drive=pressure', '--stationary',
              '--model=mrt', '--visc=0.1', '--access_pattern=AA', '--max_iters=10', '--every=5']),
    'sc3d': ('sc_separation_3d', 'SeparationSCSim', 'EqualSubdomainsGeometry3D',
             ['--lat_nx=16', '--lat_ny=12', '--lat_nz=10', '--access_pattern=AA', '--seed=7', '--max_iters=9',
              '--every=3']),
}


@pytest.mark.parametrize('name', sorted(CASES))
def test_fast_output_equals_reference_style_output(tmp_path, monkeypatch, name):
    import importlib
    import sys
    from sailfish_b200 import geo as geomod
    from sailfish_b200.controller import LBSimulationController
    modname, clsname, geoname, args = CASES[name]
    mod = importlib.import_module('sailfish_b200.examples.' + modname)
    files = {}
    for tag, extra in (('fast', []), ('sync', ['--sync_output'])):
        out = str(tmp_path / tag)
        monkeypatch.setattr(sys, 'argv', ['x'] + args + ['--output=' + out, '--quiet'] + extra)
        LBSimulationController(getattr(mod, clsname), getattr(geomod, geoname) if geoname else None).run()
        files[tag] = sorted(glob.glob(out + '.0.*.npz'))
    assert files['fast'] and len(files['fast']) == len(files['sync'])
    saw_nan = False
    for a, b in zip(files['fast'], files['sync']):
        assert a.split('.')[-2] == b.split('.')[-2]
        fa, fb = np.load(a), np.load(b)
        assert sorted(fa.files) == sorted(fb.files)
        for k in fa.files:
            assert fa[k].shape == fb[k].shape and fa[k].dtype == fb[k].dtype
            np.testing.assert_array_equal(fa[k], fb[k])          # NaN positions included
            saw_nan |= bool(np.isnan(fa[k]).any())
    if name != 'sc3d':
        assert saw_nan      # walls are masked
