"""North-star physics regtests on the GPU: native-runner versions of the reference's
regtest/ldc_3d.py (Re = 400 cavity vs the literature profiles in regtest/ldc_golden/) and
regtest/poiseuille.py (max-velocity ratio of pipe flow over a range of viscosities, compared with the
reference's own table regtest/results/poiseuille3d/D3Q19/*/force/single/fullbb.dat)."""
import json
import os

import pytest

import physics
import util

pytestmark = pytest.mark.gpu

OUT = os.path.join(util.ROOT, 'gpurun_out')


def _record(name, data):
    """Measured numbers travel back to the build container (gpurun_out/ is merged) for profiles/."""
    try:
        os.makedirs(OUT, exist_ok=True)
        path = os.path.join(OUT, 'physics_r02.json')
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = data
        json.dump(cur, open(path, 'w'), indent=1, sort_keys=True)
    except Exception:                            # noqa: BLE001  (read-only checkout: the assertions still run)
        pass


@pytest.mark.parametrize('n,steps', [(128, 50000)])
def test_ldc3d_re400_centre_lines(n, steps):
    """regtest/ldc_3d.py runs 256^3 for 100 000 steps (20 lid turn-overs n / max_v) and plots the centre-line
    profiles over the points digitised from Sheu & Tsai; here: 128^3, the same 20 turn-overs, and the
    comparison as an assertion.  Tolerance 0.04 lid velocities: the golden points were read from a graph
    (+-0.02), the grid is half the reference's, and the literature solution is a different discretisation."""
    r = physics.ldc3d_re400(n, steps)
    _record('ldc3d_re400_%d' % n, {k: v for k, v in r.items() if not k.startswith('res_')})
    assert r['finite']
    assert r['dev_vx'] < 0.04 and r['dev_vz'] < 0.04, (r['dev_vx'], r['dev_vz'])
    # shape of the Re = 400 solution: reverse flow along the bottom and a strong down-draught at the lee wall
    # (extrema of the digitised profiles: min vx -0.22, min vz -0.35, max vz +0.19)
    assert -0.30 < r['vx_min'] < -0.15, r['vx_min']
    assert -0.45 < r['vz_min'] < -0.27 and 0.12 < r['vz_max'] < 0.27, (r['vz_min'], r['vz_max'])


@pytest.mark.parametrize('model', ['bgk', 'mrt'])
@pytest.mark.parametrize('visc', [0.1, 0.05, 0.02, 0.01])
def test_poiseuille3d_max_velocity(model, visc):
    """|max(v)/max(v_th) - 1|: within 1 % of theory and within 5e-3 of what the reference itself measured at
    this viscosity (its table spans -0.3 % .. -0.9 % for visc >= 0.01)."""
    r = physics.poiseuille3d_max_velocity_error(visc, model, width=64, length=32)
    ref = physics.reference_poiseuille_error(visc, model)
    _record('poiseuille3d_%s_%g' % (model, visc), dict(r, reference=ref))
    assert r['finite']
    assert abs(r['err']) < 1e-2, r
    assert abs(r['err'] - ref) < 5e-3, (r, ref)


@pytest.mark.parametrize('direction', ['y', 'z'])
def test_poiseuille3d_other_axes(direction):
    """The same pipe along y and z (z = the split axis of the native runner) gives the same error."""
    a = physics.poiseuille3d_max_velocity_error(0.05, 'bgk', width=48, length=16, direction='x')
    b = physics.poiseuille3d_max_velocity_error(0.05, 'bgk', width=48, length=16, direction=direction)
    assert abs(a['err'] - b['err']) < 2e-5, (a, b)
