#!/usr/bin/env python
"""2D lid-driven cavity: user code of the reference's examples/ldc_2d.py:1-41 (BASELINE config 1)."""
from sailfish_b200.subdomain import Subdomain2D
from sailfish_b200.node_type import NTFullBBWall, NTRegularizedVelocity
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.lb_single import LBFluidSim


class LDCBlock(Subdomain2D):
    max_v = 0.1

    def boundary_conditions(self, hx, hy):
        wall_bc = NTFullBBWall
        velocity_bc = NTRegularizedVelocity
        wall_map = (hx == self.gx - 1) | (hx == 0) | (hy == 0)
        self.set_node((hy == self.gy - 1) & (hx > 0) & (hx < self.gx - 1), velocity_bc((self.max_v, 0.0)))
        self.set_node(wall_map, wall_bc)

    def initial_conditions(self, sim, hx, hy):
        sim.rho[:] = 1.0
        sim.vx[hy == self.gy - 1] = self.max_v


class LDCSim(LBFluidSim):
    subdomain = LDCBlock

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update({'lat_nx': 256, 'lat_ny': 256})


if __name__ == '__main__':
    LBSimulationController(LDCSim).run()
