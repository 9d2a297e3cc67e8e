"""Binary Shan-Chen with the domain split along z (periodic wrap included): populations AND the
rho / phi ghost layers travel between subdomains; the result must equal the single-subdomain run
bit for bit (the reference's regtest/subdomains/binary_pbc.py bar is 6 decimals)."""
import numpy as np
import pytest

from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.examples import sc_separation_3d

pytestmark = pytest.mark.gpu


def run(args, steps):
    from sailfish_b200.controller import LBSimulationController
    ctrl = LBSimulationController(sc_separation_3d.SeparationSCSim, EqualSubdomainsGeometry3D)
    ctrl.setup(args + ['--quiet', '--seed=77'])
    # identical initial state regardless of the decomposition: overwrite with a deterministic field
    for r in ctrl.runners:
        sim, spec = r._sim, r._spec
        hz, hy, hx = np.meshgrid(np.arange(spec.oz, spec.oz + spec.nz), np.arange(spec.ny), np.arange(spec.nx),
                                 indexing='ij')
        sim.rho[:] = 1.0 + 1e-3 * np.sin(0.7 * hx + 1.3 * hy + 0.4 * hz)
        sim.phi[:] = 1.0 + 1e-3 * np.cos(0.5 * hx - 0.9 * hy + 1.1 * hz)
        dev = r._dev
        for b in (dev.d_rho, dev.d_phi):
            dev.be.to_buf(b)
        dev.set_initial_conditions()
    for k in range(steps):
        ctrl.step_all(k == steps - 1)
    ctrl.sync()
    pieces = []
    for r in ctrl.runners:
        r._fields_to_host()
        sl = r._spec._nonghost_slice
        pieces.append((r._spec.location[2], [np.array(r.field_base(r._sim.rho))[sl],
                                             np.array(r.field_base(r._sim.phi))[sl]]))
    pieces.sort(key=lambda t: t[0])
    return [np.concatenate([p[1][k] for p in pieces], axis=0) for k in range(2)]


@pytest.mark.parametrize('ap', ['AB', 'AA'])
@pytest.mark.parametrize('n', [2, 3])
def test_split_equals_single(ap, n):
    args = ['--lat_nx=20', '--lat_ny=18', '--lat_nz=24', '--access_pattern=' + ap]
    one = run(args + ['--subdomains=1'], 10)
    many = run(args + ['--subdomains=%d' % n], 10)
    for a, b in zip(one, many):
        assert np.isfinite(a).all()
        assert np.array_equal(a, b), np.abs(a - b).max()
