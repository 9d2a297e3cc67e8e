#!/usr/bin/env python
"""Poiseuille flow through a circular pipe (BASELINE config 3; same set-up as the reference's
examples/poiseuille_3d.py): full-way bounce-back pipe wall; the flow is driven either by a density
difference between equilibrium-density inlet and outlet planes (--drive=pressure) or by a Guo body
force with the pipe axis periodic (--drive=force)."""
import numpy as np

from sailfish_b200.controller import LBSimulationController
from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.lb_base import LBForcedSim
from sailfish_b200.lb_single import LBFluidSim
from sailfish_b200.node_type import NTEquilibriumDensity, NTFullBBWall
from sailfish_b200.subdomain import Subdomain3D

AXES = 'xyz'


def _lat_sizes(config):
    return dict(x=config.lat_nx, y=config.lat_ny, z=config.lat_nz)


class Pipe(Subdomain3D):
    max_v = 0.02                 # centre-line velocity of the analytic profile
    wall_bc = NTFullBBWall

    # -- analytic relations -------------------------------------------------------------------------
    @classmethod
    def width(cls, config):
        """Edge of the square cross-section the pipe is inscribed in."""
        sizes = _lat_sizes(config)
        return min(sizes[a] for a in AXES if a != config.flow_direction)

    @classmethod
    def channel_width(cls, config):
        """Hydraulic diameter: the wall sits `location` nodes inside the first / last wall node."""
        return cls.width(config) - 1 - 2 * cls.wall_bc.location

    @classmethod
    def driving_gradient(cls, config):
        """dp/dx (= body acceleration for rho = 1) of Hagen-Poiseuille flow with centre velocity max_v."""
        return cls.max_v * (16.0 * config.visc) / cls.channel_width(config) ** 2

    @property
    def pressure_delta(self):
        length = _lat_sizes(self.config)[self.config.flow_direction] - 1
        # (same association of the products as the analytic expression in the reference example, so
        #  that the imposed densities agree to the last bit)
        return self.max_v * (16.0 * self.config.visc) * length / self.channel_width(self.config) ** 2

    # -- geometry -----------------------------------------------------------------------------------
    def _along(self, hx, hy, hz):
        """Coordinate array and extent along the pipe axis."""
        return dict(x=(hx, self.gx), y=(hy, self.gy), z=(hz, self.gz))[self.config.flow_direction]

    def _outside_pipe(self, hx, hy, hz):
        coords = dict(x=(hx, self.gx), y=(hy, self.gy), z=(hz, self.gz))
        (a, na), (b, nb) = [coords[ax] for ax in AXES if ax != self.config.flow_direction]
        radius_sq = (self.channel_width(self.config) / 2.0) ** 2
        return (a - (na / 2 - 0.5)) ** 2 + (b - (nb / 2 - 0.5)) ** 2 >= radius_sq

    def boundary_conditions(self, hx, hy, hz):
        solid = self._outside_pipe(hx, hy, hz)
        self.set_node(solid, self.wall_bc)
        if self.config.drive != 'pressure':
            return
        h, n = self._along(hx, hy, hz)
        open_section = np.logical_not(solid)
        half = 3.0 * self.pressure_delta / 2.0
        self.set_node((h == 0) & open_section, NTEquilibriumDensity(1.0 + half))
        self.set_node((h == n - 1) & open_section, NTEquilibriumDensity(1.0 - half))

    def initial_conditions(self, sim, hx, hy, hz):
        sim.rho[:] = 1.0
        if self.config.stationary and self.config.drive == 'pressure':
            h, n = self._along(hx, hy, hz)                 # start from the linear pressure drop
            sim.rho[:] = 1.0 + 3.0 * self.pressure_delta * (n / 2.0 - h)


class PoiseuilleSim(LBFluidSim, LBForcedSim):
    subdomain = Pipe

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update(dict(grid='D3Q19', lat_nx=64, lat_ny=64, lat_nz=64, visc=0.1))

    @classmethod
    def add_options(cls, group, dim):
        group.add_argument('--flow_direction', choices=list(AXES), default='x', type=str)
        group.add_argument('--drive', choices=['force', 'pressure'], default='force', type=str)
        group.add_argument('--stationary', action='store_true', default=False)

    @classmethod
    def modify_config(cls, config):
        if config.drive == 'force':        # the pipe closes on itself along its axis
            for axis in AXES:
                setattr(config, 'periodic_' + axis, axis == config.flow_direction)

    def __init__(self, config):
        super(PoiseuilleSim, self).__init__(config)
        if config.drive == 'force':
            g = self.subdomain.driving_gradient(config)
            self.add_body_force(tuple(g if axis == config.flow_direction else 0.0 for axis in AXES))


if __name__ == '__main__':
    LBSimulationController(PoiseuilleSim, EqualSubdomainsGeometry3D).run()
