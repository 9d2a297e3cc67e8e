"""Sweeps of the native runner's decompositions on the CPU stand-in device: model x body force x locally periodic axes x
split, single fluid and binary Shan-Chen, 2-D and 3-D, stepped and replayed.  Every split run must give the
single-subdomain fields bit for bit (the reference's bar is 6 decimals, regtest/subdomains/util.py:27-30).  These
sweeps found the 2-D periodic-row and the D3Q27 landing-zone defects of the linked periodic kernel."""
import numpy as np
import pytest

import block_util
import fake_device
from sailfish_b200 import node_type as nt
from sailfish_b200.geo import (BlockSubdomainsGeometry3D, EqualSubdomainsGeometry2D, EqualSubdomainsGeometry3D)
from sailfish_b200.lb_base import LBForcedSim
from sailfish_b200.lb_binary import LBBinaryFluidShanChen
from sailfish_b200.lb_single import LBFluidSim
from sailfish_b200.subdomain import Subdomain2D, Subdomain3D


@pytest.fixture
def fake(oracle, monkeypatch):
    return fake_device.install_fake(monkeypatch)


def _walls(sd, walls, h):
    w = np.zeros(np.broadcast(*h).shape, bool)
    for a, c, g in zip('xyz', h, (sd.gx, sd.gy, getattr(sd, 'gz', 0))):
        if a in walls:
            w |= (c == 0) | (c == g - 1)
    if w.any():
        sd.set_node(w, nt.NTFullBBWall)


def _single_fluid(dim, per, walls, grid, force):
    class Sub(Subdomain3D if dim == 3 else Subdomain2D):
        def boundary_conditions(self, *h):
            _walls(self, walls, h)

        def initial_conditions(self, sim, *h):
            sim.rho[:] = 1.0 + 1e-3 * np.sin(2 * np.pi * h[0] / self.gx)
            sim.vx[:] = 0.04 * np.sin(2 * np.pi * h[1] / self.gy)
            sim.vy[:] = 0.02 * np.cos(2 * np.pi * h[0] / self.gx)

    class Sim(LBFluidSim, LBForcedSim):
        subdomain = Sub

        def __init__(self, config):
            super(Sim, self).__init__(config)
            if force:
                self.add_body_force((1e-4, -2e-4, 3e-4)[:dim])

        @classmethod
        def update_defaults(cls, d):
            d.update(lat_nx=10, lat_ny=9, grid=grid, **dict(('periodic_' + a, True) for a in per))
            if dim == 3:
                d.update(lat_nz=12)
    return Sim


def _binary(dim, per, walls):
    class Sub(Subdomain3D if dim == 3 else Subdomain2D):
        def boundary_conditions(self, *h):
            _walls(self, walls, h)

        def initial_conditions(self, sim, *h):
            sim.rho[:] = 1.0 + 0.05 * np.sin(2 * np.pi * h[0] / self.gx) * np.cos(2 * np.pi * h[1] / self.gy)
            sim.phi[:] = 1.0 + 0.05 * np.cos(2 * np.pi * h[-1] / (self.gz if dim == 3 else self.gy))

    class Sim(LBBinaryFluidShanChen):
        subdomain = Sub

        @classmethod
        def update_defaults(cls, d):
            d.update(lat_nx=10, lat_ny=9, grid='D3Q19' if dim == 3 else 'D2Q9', G12=1.2, visc=1.0 / 6.0, tau_phi=1.0,
                     **dict(('periodic_' + a, True) for a in per))
            if dim == 3:
                d.update(lat_nz=12)
    return Sim


def _same(sim, geo, args, one, split, gsize, steps=8, replay=False, fields=block_util.collect):
    def go(extra, rp):
        from sailfish_b200.controller import LBSimulationController
        ctrl = LBSimulationController(sim, geo)
        ctrl.setup(args + extra + ['--quiet'])
        if rp:
            ctrl.run_steps(steps, save_last=True)
        else:
            for k in range(steps):
                ctrl.step_all(k == steps - 1)
        ctrl.sync()
        return fields(ctrl)
    np.testing.assert_array_equal(block_util.merge(go(one, False), gsize), block_util.merge(go(split, replay), gsize))


PER3 = [('xy', 'z'), ('x', 'yz'), ('', 'xyz')]


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('model', ['bgk', 'mrt'])
@pytest.mark.parametrize('force', [False, True], ids=['noforce', 'force'])
@pytest.mark.parametrize('per,walls', PER3, ids=[p or 'none' for p, _ in PER3])
def test_single_fluid_3d(fake, per, walls, force, model, ap):
    sim = _single_fluid(3, per, walls, 'D3Q19', force)
    args = ['--access_pattern=' + ap, '--model=' + model, '--visc=0.05']
    _same(sim, EqualSubdomainsGeometry3D, args, ['--subdomains=1'], ['--subdomains=3'], (10, 9, 12))
    _same(sim, EqualSubdomainsGeometry3D, args, ['--subdomains=1'], ['--subdomains=2'], (10, 9, 12), replay=True)
    if not per:
        _same(sim, BlockSubdomainsGeometry3D, args, ['--blocks=1,1,1'], ['--blocks=2,2,2'], (10, 9, 12))


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('model', ['bgk', 'mrt'])
@pytest.mark.parametrize('per,walls', [('x', 'y'), ('', 'xy')], ids=['x', 'none'])
def test_single_fluid_2d(fake, per, walls, model, ap):
    sim = _single_fluid(2, per, walls, 'D2Q9', True)
    args = ['--access_pattern=' + ap, '--model=' + model, '--visc=0.05']
    _same(sim, EqualSubdomainsGeometry2D, args, ['--subdomains=1'], ['--subdomains=3'], (10, 9))


def _collect_binary(ctrl):
    out = []
    for r in ctrl.runners:
        r._fields_to_host()
        sl = r._spec._nonghost_slice
        s = r._sim
        f = [np.array(r.field_base(s.rho))[sl], np.array(r.field_base(s.phi))[sl]] + [np.array(r.field_base(c))[sl] for c in s.v]
        out.append((tuple(r._spec.location), tuple(r._spec.size), np.array(f)))
    return out


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('per,walls', [('xyz', ''), ('xy', 'z'), ('x', 'yz'), ('', 'xyz'), ('z', 'xy')],
                         ids=['xyz', 'xy', 'x', 'none', 'z'])
def test_binary_shan_chen_3d(fake, per, walls, ap):
    sim = _binary(3, per, walls)
    args = ['--access_pattern=' + ap]
    _same(sim, EqualSubdomainsGeometry3D, args, ['--subdomains=1'], ['--subdomains=3'], (10, 9, 12), fields=_collect_binary)
    _same(sim, EqualSubdomainsGeometry3D, args, ['--subdomains=1'], ['--subdomains=2'], (10, 9, 12), replay=True,
          fields=_collect_binary)


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('per,walls', [('xy', ''), ('x', 'y'), ('', 'xy')], ids=['xy', 'x', 'none'])
def test_binary_shan_chen_2d(fake, per, walls, ap):
    sim = _binary(2, per, walls)
    _same(sim, EqualSubdomainsGeometry2D, ['--access_pattern=' + ap], ['--subdomains=1'], ['--subdomains=3'], (10, 9),
          fields=_collect_binary)
