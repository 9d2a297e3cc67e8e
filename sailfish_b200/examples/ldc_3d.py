#!/usr/bin/env python
"""3D lid-driven cavity: the user code of the reference's examples/ldc_3d.py:1-43, unchanged
apart from the package name (BASELINE config 2 / 4)."""
import numpy as np
from sailfish_b200.subdomain import Subdomain3D
from sailfish_b200.node_type import NTFullBBWall, NTRegularizedVelocity
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.lb_single import LBFluidSim


class LDCBlock(Subdomain3D):
    max_v = 0.05

    def boundary_conditions(self, hx, hy, hz):
        wall_bc = NTFullBBWall
        velocity_bc = NTRegularizedVelocity
        wall_map = ((hz == 0) | (hx == self.gx - 1) | (hx == 0) | (hy == 0) | (hy == self.gy - 1))
        self.set_node(wall_map, wall_bc)
        self.set_node((hz == self.gz - 1) & np.logical_not(wall_map), velocity_bc((self.max_v, 0.0, 0.0)))

    def initial_conditions(self, sim, hx, hy, hz):
        sim.rho[:] = 1.0
        sim.vx[hz == self.gz - 1] = self.max_v


class LDCSim(LBFluidSim):
    subdomain = LDCBlock

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update({'lat_nx': 64, 'lat_ny': 64, 'lat_nz': 64, 'grid': 'D3Q19'})


if __name__ == '__main__':
    LBSimulationController(LDCSim).run()
