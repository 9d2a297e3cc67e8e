#!/usr/bin/env python
"""Binary Shan-Chen phase separation: user code of the reference's
examples/binary_fluid/sc_separation_3d.py:1-33 (BASELINE config 5)."""
import numpy as np

from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.subdomain import Subdomain3D
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.lb_binary import LBBinaryFluidShanChen


class SeparationDomain(Subdomain3D):
    def initial_conditions(self, sim, hx, hy, hz):
        sim.rho[:] = 1.0 + np.random.rand(*sim.rho.shape) / 1000.0
        sim.phi[:] = 1.0 + np.random.rand(*sim.phi.shape) / 1000.0

    def boundary_conditions(self, hx, hy, hz):
        pass


class SeparationSCSim(LBBinaryFluidShanChen):
    subdomain = SeparationDomain

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update({'lat_nx': 192, 'lat_ny': 192, 'lat_nz': 192, 'grid': 'D3Q19', 'G12': 1.2,
                         'visc': 1.0 / 6.0, 'periodic_x': True, 'periodic_y': True, 'periodic_z': True})


if __name__ == '__main__':
    LBSimulationController(SeparationSCSim, EqualSubdomainsGeometry3D).run()
