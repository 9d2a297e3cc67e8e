"""Stepping from C (slf_run_steps: one cooperative multi-step launch for small lattices without periodic axes,
CUDA-graph replay of step pairs otherwise) must give bit for bit what one CollideAndPropagate launch per step
gives -- populations and macroscopic fields, odd and even step counts, with and without the final save."""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

CASES = ['ldc2d_bgk_AA', 'ldc2d_mrt_AA', 'ldc3d_bgk_AA', 'ldc3d_mrt_AA', 'ldc3d_q27_AA', 'poi3d_press_mrt_AA',
         'zoo_solid_box2d_AA', 'zoo_pressure_box3d_AA', 'zoo_channel2d_AA',
         'box3d_bgk_AA', 'poi3d_force_bgk_AA', 'ldc3d_bgk_AB']          # the last three replay graphs (PBC / AB)


def _run(case, steps, stepper):
    from sailfish_b200.runtime import DeviceLattice
    c = util.load_case(case)
    p = util.params_from_case(c)
    dev = DeviceLattice(p, c['map'], c['rho0'], c['v0'])
    getattr(dev, stepper)(steps, save_last=True)
    rho, v = dev.fields_to_host()
    return dev.get_dist(), np.array(rho), [np.array(x) for x in v]


def _same(a, b):
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.mark.parametrize('case', CASES)
@pytest.mark.parametrize('steps', [2, 7, 40])
def test_run_steps_equals_one_launch_per_step(case, steps):
    if case not in util.ZOO_STEP_CASES and not case.startswith(tuple(util.CASES)):
        pytest.skip('no such golden case')
    a = _run(case, steps, 'run')
    b = _run(case, steps, 'run_python_loop')
    assert _same(a[0], b[0]), 'populations differ'
    assert _same(a[1], b[1]), 'rho differs'
    for x, y in zip(a[2], b[2]):
        assert _same(x, y), 'v differs'


def test_multistep_launch_is_taken_for_small_wall_bounded_lattices():
    """ldc_2d: after run(100) the library has counted ONE launch (the cooperative kernel), not 200."""
    from sailfish_b200 import backend_b200 as bk
    from sailfish_b200.runtime import DeviceLattice
    c = util.load_case('ldc2d_bgk_AA')
    dev = DeviceLattice(util.params_from_case(c), c['map'], c['rho0'], c['v0'])
    dev.sync()
    n0 = bk.lib().launch_count()
    dev.run(100, save_last=False)
    dev.sync()
    assert bk.lib().launch_count() - n0 == 1
