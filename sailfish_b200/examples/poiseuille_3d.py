#!/usr/bin/env python
"""3D Poiseuille flow in a pipe: user code of the reference's examples/poiseuille_3d.py:14-200
(BASELINE config 3): full-way bounce-back pipe wall, equilibrium-density inlet/outlet
(--drive=pressure) or Guo body force with a periodic axis (--drive=force)."""
import numpy as np

from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.subdomain import Subdomain3D
from sailfish_b200.node_type import NTFullBBWall, NTEquilibriumDensity
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.lb_single import LBFluidSim
from sailfish_b200.lb_base import LBForcedSim


class PoiseuilleSubdomain(Subdomain3D):
    max_v = 0.02
    wall_bc = NTFullBBWall

    @property
    def channel_length(self):
        return {'x': self.config.lat_nx, 'y': self.config.lat_ny, 'z': self.config.lat_nz}[
            self.config.flow_direction] - 1

    @property
    def pressure_delta(self):
        return (self.max_v * (16.0 * self.config.visc) * self.channel_length /
                (self.channel_width(self.config) ** 2))

    def _set_pressure_bc(self, hx, hy, hz, wall_map):
        not_wall = np.logical_not(wall_map)
        h, n = {'x': (hx, self.gx), 'y': (hy, self.gy), 'z': (hz, self.gz)}[self.config.flow_direction]
        inlet_map = (h == 0) & not_wall
        outlet_map = (h == n - 1) & not_wall
        pressure = self.pressure_delta
        self.set_node(inlet_map, NTEquilibriumDensity(1.0 + 3.0 * pressure / 2.0))
        self.set_node(outlet_map, NTEquilibriumDensity(1.0 - 3.0 * pressure / 2.0))

    def boundary_conditions(self, hx, hy, hz):
        radiussq = (self.channel_width(self.config) / 2.0) ** 2
        fd = self.config.flow_direction
        if fd == 'z':
            wall_map = (hx - (self.gx / 2 - 0.5)) ** 2 + (hy - (self.gy / 2 - 0.5)) ** 2 >= radiussq
        elif fd == 'y':
            wall_map = (hx - (self.gx / 2 - 0.5)) ** 2 + (hz - (self.gz / 2 - 0.5)) ** 2 >= radiussq
        else:
            wall_map = (hy - (self.gy / 2 - 0.5)) ** 2 + (hz - (self.gz / 2 - 0.5)) ** 2 >= radiussq
        self.set_node(wall_map, self.wall_bc)
        if self.config.drive == 'pressure':
            self._set_pressure_bc(hx, hy, hz, wall_map)

    def initial_conditions(self, sim, hx, hy, hz):
        sim.rho[:] = 1.0
        if not self.config.stationary:
            return
        if self.config.drive == 'pressure':
            pressure = self.pressure_delta
            h, n = {'x': (hx, self.gx), 'y': (hy, self.gy), 'z': (hz, self.gz)}[self.config.flow_direction]
            sim.rho[:] = 1.0 + 3.0 * pressure * (n / 2.0 - h)

    @classmethod
    def channel_width(cls, config):
        return cls.width(config) - 1 - 2 * cls.wall_bc.location

    @classmethod
    def width(cls, config):
        if config.flow_direction == 'x':
            return min(config.lat_ny, config.lat_nz)
        elif config.flow_direction == 'y':
            return min(config.lat_nx, config.lat_nz)
        return min(config.lat_nx, config.lat_ny)


class PoiseuilleSim(LBFluidSim, LBForcedSim):
    subdomain = PoiseuilleSubdomain

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update({'lat_nx': 64, 'lat_ny': 64, 'lat_nz': 64, 'visc': 0.1, 'grid': 'D3Q19'})

    @classmethod
    def add_options(cls, group, dim):
        group.add_argument('--flow_direction', type=str, default='x', choices=['x', 'y', 'z'])
        group.add_argument('--stationary', action='store_true', default=False)
        group.add_argument('--drive', type=str, default='force', choices=['force', 'pressure'])

    @classmethod
    def modify_config(cls, config):
        if config.drive == 'force':
            config.periodic_x = config.flow_direction == 'x'
            config.periodic_y = config.flow_direction == 'y'
            config.periodic_z = config.flow_direction == 'z'

    def __init__(self, config):
        super(PoiseuilleSim, self).__init__(config)
        if config.drive == 'force':
            cw = self.subdomain.channel_width(config)
            accel = self.subdomain.max_v * (16.0 * config.visc) / cw ** 2
            self.add_body_force({'x': (accel, 0.0, 0.0), 'y': (0.0, accel, 0.0), 'z': (0.0, 0.0, accel)}[
                config.flow_direction])


if __name__ == '__main__':
    LBSimulationController(PoiseuilleSim, EqualSubdomainsGeometry3D).run()
