"""Model base classes: mirror of sailfish/lb_base.py (LBSim :30-302, LBForcedSim :305-394)
for the single-fluid hot path."""
from collections import namedtuple

import numpy as np

from sailfish_b200 import lattice

ScalarField = namedtuple('ScalarField', 'name')
VectorField = namedtuple('VectorField', 'name')
# What the reference hands to code generation under ctx['forces'] (lb_base.py:316, :363): numeric forces as
# {grid: {is_acceleration: vector}}, symbolic ones by expression.
ForcePair = namedtuple('ForcePair', 'numeric symbolic')


class LBSim(object):
    subdomain = None
    subdomain_runner = None
    nonlocality = 0

    @classmethod
    def add_options(cls, group, dim):
        """lb_base.py:54-100 (options that reach the kernels)."""
        grids = [n for n in lattice.KNOWN if lattice.get(n).dim == dim]
        group.add_argument('--grid', help='LB grid', type=str, choices=grids, default=grids[0])
        group.add_argument('--access_pattern', type=str, default='AB', choices=('AB', 'AA'),
                           help='AB: two lattice copies; AA: single copy, in-place updates')
        group.add_argument('--node_addressing', type=str, default='direct', choices=('direct',),
                           help='only direct (dense) addressing is on the B200 hot path')
        group.add_argument('--nouse_link_tags', action='store_false', dest='use_link_tags', default=True)

    @classmethod
    def modify_config(cls, config):
        pass

    @classmethod
    def update_defaults(cls, defaults):
        pass

    def __init__(self, config):
        self.config = config
        self.iteration = 0
        self.need_sync_flag = False
        self.need_fields_flag = False
        self.grid = lattice.get(config.grid)
        self.grids = [self.grid]

    @property
    def dim(self):
        return self.grid.dim

    @classmethod
    def fields(cls):
        return []

    def constants(self):
        """lb_base.py:104-107: named constants of the model (the Shan-Chen couplings)."""
        return {}

    def update_context(self, ctx):
        """The model's share of the kernel configuration, under the keys the reference passes to code
        generation (lb_base.py:120-137).  `params.from_context` turns the finished dict into SlfParams; the
        same function serves the unmodified reference on --backends=b200, so both ways in configure the
        kernels through one path."""
        cfg = self.config
        ctx.update(grid=self.grid, grids=self.grids, constants=self.constants(), config=cfg,
                   access_pattern=cfg.access_pattern)
        for key, default in (('relaxation_enabled', True), ('propagation_enabled', True),
                             ('node_addressing', 'direct'), ('propagate_on_read', False)):
            ctx[key] = getattr(cfg, key, default)

    def get_state(self):
        """lb_base.py:216-220: what a checkpoint stores besides the populations."""
        return {'iteration': self.iteration}

    def set_state(self, state):
        self.iteration = state['iteration']

    def need_checkpoint(self):
        """lb_base.py:254-260."""
        cfg = self.config
        every = getattr(cfg, 'checkpoint_every', 0)
        return every > 0 and (self.iteration % every) == 0 and self.iteration >= getattr(cfg, 'checkpoint_from', 0)

    def need_output(self):
        """lb_base.py:222-231."""
        if self.config.output_required:
            return ((self.iteration + 1) % self.config.every) == 0 and self.config.from_ <= self.iteration
        return False

    def need_sync_fields(self):
        """lb_base.py:233-252."""
        need_sync = self.need_sync_flag or self.need_output()
        need_fields = self.need_fields_flag or need_sync
        self.need_fields_flag = False
        self.need_sync_flag = False
        return need_sync, need_fields

    def before_main_loop(self, runner):
        pass

    def after_step(self, runner):
        pass

    def after_main_loop(self, runner):
        pass

    def body_acceleration(self):
        return None


class LBForcedSim(LBSim):
    """Body forces (lb_base.py:305-394).  Constant accelerations with the Guo scheme are on the hot path: on the
    single lattice (force-driven Poiseuille) and on either component of a binary mixture."""

    def __init__(self, config):
        super(LBForcedSim, self).__init__(config)
        self._forces = {}
        self._force_couplings = {}

    @classmethod
    def add_options(cls, group, dim):
        group.add_argument('--force_implementation', type=str, choices=['guo'], default='guo')

    def add_body_force(self, force, grid=0, accel=True):
        dim = self.grid.dim
        assert len(force) == dim
        if not accel:
            raise ValueError('B200 hot path: acceleration-type body forces only (accel=True)')
        if grid < 0 or grid >= len(self.grids):
            raise ValueError('no lattice %d in this model' % grid)
        cur = self._forces.setdefault(grid, np.zeros(dim, np.float64))
        self._forces[grid] = cur + np.float64(force)

    def update_context(self, ctx):
        """lb_base.py:361-366."""
        super(LBForcedSim, self).update_context(ctx)
        numeric = dict((g, {True: tuple(float(x) for x in a)}) for g, a in self._forces.items())
        ctx.update(forces=ForcePair(numeric, {}), force_couplings=dict(self._force_couplings),
                   force_implementation=getattr(self.config, 'force_implementation', 'guo'))

    def add_force_coupling(self, grid_a, grid_b, const_name):
        """lb_base.py:386-394: Shan-Chen interaction of component grid_a with grid_b, strength
        constants()[const_name]."""
        self._force_couplings[(grid_a, grid_b)] = const_name

    def body_acceleration(self, grid=0):
        if grid not in self._forces:
            return None
        a = self._forces[grid]
        return [float(x) for x in a] + [0.0] * (3 - len(a))
