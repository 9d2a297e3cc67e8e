"""Every executable example of the reference (what its tests/run_examples.sh runs), UNCHANGED, for four
steps on the b200 backend.  Pins which ones the hot path covers -- they must run and write their output --
and that everything else is refused with UnsupportedConfiguration naming the feature (never a silent wrong
answer, never a crash inside a kernel).  Device = tests/fake_device.py (CPU oracle)."""
import os
import subprocess
import sys

import pytest

import util

REF = os.environ.get('SAILFISH_REFERENCE', '/root/reference')
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'examples')), reason='reference tree not available')

RUNS = ['cylinder', 'duct_flow', 'external_geometry', 'fs_gaussian', 'ldc_2d', 'ldc_3d', 'poiseuille', 'sphere_3d',
        'taylor_green_2d', 'binary_fluid/sc_capillary_wave_2d', 'binary_fluid/sc_laplace_2d',
        'binary_fluid/sc_separation_2d', 'binary_fluid/sc_separation_3d', 'binary_fluid/sc_separation_3d_walls',
        'binary_fluid/sc_capillary', 'binary_fluid/sc_rayleigh_taylor_2d']
# example -> fragment of the refusal message
REFUSED = {
    'binary_fluid/binary_microchannel': 'free-energy', 'binary_fluid/fe_capillary_wave_2d': 'free-energy',
    'binary_fluid/fe_poiseuille_2d': 'free-energy', 'binary_fluid/fe_separation_2d': 'free-energy',
    'binary_fluid/fe_separation_3d': 'free-energy', 'binary_fluid/fe_viscous_fingering': 'free-energy',
    'binary_fluid/sc_poiseuille_2d': 'force_implementation=guo', 'four_rolls_mill': 'symbolic body forces',
    'ldc_2d_entropic': 'relaxation model elbm', 'poiseuille_3d': 'lattice D3Q13',     # its default grid
    'poiseuille_pulsatile': 'time/space dependent', 'sc_drop': 'single-component Shan-Chen',
    'sc_phase_separation': 'single-component Shan-Chen', 'turbulence/kida_vortex': 'minimize_roundoff',
}
# not run: ldc_2d_unorm (broken import in the reference itself), ternary_fluid/sc_drop_2d (matplotlib),
# turbulence/channel_flow and channel_cube (netifaces)


@pytest.fixture(scope='module')
def verdicts(oracle, tmp_path_factory):
    out = tmp_path_factory.mktemp('examples')
    names = RUNS + sorted(REFUSED)
    res, running = {}, []

    def reap(block):
        for name, p in list(running):
            if block or p.poll() is not None:
                so, se = p.communicate(timeout=600)
                res[name] = (so, se)
                running.remove((name, p))
                if block:
                    return

    for name in names:
        cmd = [sys.executable, os.path.join(util.ROOT, 'tests', 'reference_example_runner.py'), REF,
               'examples/%s.py' % name, str(out / name.replace('/', '_'))]
        running.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
        while len(running) >= (os.cpu_count() or 4):
            reap(True)
    while running:
        reap(True)
    return res


@pytest.mark.parametrize('name', RUNS)
def test_example_runs_on_b200_backend(verdicts, name):
    so, se = verdicts[name]
    assert 'VERDICT ran' in so, se[-2000:]


@pytest.mark.parametrize('name', sorted(REFUSED))
def test_example_outside_the_hot_path_is_refused_by_name(verdicts, name):
    so, se = verdicts[name]
    assert 'VERDICT no output' in so
    assert 'UnsupportedConfiguration' in se and REFUSED[name] in se, se[-2000:]
