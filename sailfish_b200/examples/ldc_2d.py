#!/usr/bin/env python
"""Lid-driven cavity in a square (BASELINE config 1; same set-up as the reference's examples/ldc_2d.py):
the top row slides along +x at 0.1 between two wall corners."""
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.lb_single import LBFluidSim
from sailfish_b200.node_type import NTFullBBWall, NTRegularizedVelocity
from sailfish_b200.subdomain import Subdomain2D


class Cavity(Subdomain2D):
    lid_speed = 0.1

    def boundary_conditions(self, hx, hy):
        top = hy == self.gy - 1
        interior_columns = (hx > 0) & (hx < self.gx - 1)
        self.set_node(top & interior_columns, NTRegularizedVelocity((self.lid_speed, 0.0)))
        self.set_node((hy == 0) | (hx == 0) | (hx == self.gx - 1), NTFullBBWall)

    def initial_conditions(self, sim, hx, hy):
        sim.rho[:] = 1.0
        sim.vx[hy == self.gy - 1] = self.lid_speed


class LDCSim(LBFluidSim):
    subdomain = Cavity

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update(dict(lat_nx=256, lat_ny=256))


if __name__ == '__main__':
    LBSimulationController(LDCSim).run()
