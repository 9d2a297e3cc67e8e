#!/usr/bin/env python
"""Spinodal decomposition of a binary Shan-Chen mixture in a periodic box (BASELINE config 5; same
set-up as the reference's examples/binary_fluid/sc_separation_3d.py): both densities start at 1 plus
uniform noise of amplitude 1e-3 drawn from numpy's global generator (seeded by --seed)."""
import numpy as np

from sailfish_b200.controller import LBSimulationController
from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.lb_binary import LBBinaryFluidShanChen
from sailfish_b200.subdomain import Subdomain3D

NOISE = 1e-3


class PeriodicBox(Subdomain3D):
    def boundary_conditions(self, hx, hy, hz):
        """Nothing but fluid."""

    def initial_conditions(self, sim, hx, hy, hz):
        for field in (sim.rho, sim.phi):        # order matters: rho consumes the random stream first
            field[:] = 1.0 + np.random.rand(*field.shape) / (1.0 / NOISE)


class SeparationSCSim(LBBinaryFluidShanChen):
    subdomain = PeriodicBox

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update(dict(grid='D3Q19', lat_nx=192, lat_ny=192, lat_nz=192, visc=1.0 / 6.0, G12=1.2))
        defaults.update(dict(('periodic_' + axis, True) for axis in 'xyz'))


if __name__ == '__main__':
    LBSimulationController(SeparationSCSim, EqualSubdomainsGeometry3D).run()
