#!/usr/bin/env python
"""Lid-driven cavity in a cube (BASELINE configs 2 and 4; same set-up as the reference's
examples/ldc_3d.py): five resting walls, the z = top plane slides along +x at 0.05."""
import numpy as np

from sailfish_b200.controller import LBSimulationController
from sailfish_b200.lb_single import LBFluidSim
from sailfish_b200.node_type import NTFullBBWall, NTRegularizedVelocity
from sailfish_b200.subdomain import Subdomain3D


class Cavity(Subdomain3D):
    lid_speed = 0.05

    def _resting_walls(self, hx, hy, hz):
        sides = (hx == 0) | (hx == self.gx - 1) | (hy == 0) | (hy == self.gy - 1)
        return sides | (hz == 0)

    def boundary_conditions(self, hx, hy, hz):
        walls = self._resting_walls(hx, hy, hz)
        lid = (hz == self.gz - 1) & np.logical_not(walls)
        self.set_node(walls, NTFullBBWall)
        self.set_node(lid, NTRegularizedVelocity((self.lid_speed, 0.0, 0.0)))

    def initial_conditions(self, sim, hx, hy, hz):
        sim.rho[:] = 1.0
        sim.vx[hz == self.gz - 1] = self.lid_speed


class LDCSim(LBFluidSim):
    subdomain = Cavity

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update(dict(grid='D3Q19', lat_nx=64, lat_ny=64, lat_nz=64))


if __name__ == '__main__':
    LBSimulationController(LDCSim).run()
