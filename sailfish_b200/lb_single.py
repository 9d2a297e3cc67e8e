"""Single-fluid model class: mirror of sailfish/lb_single.py:14-199 (LBFluidSim)."""
from sailfish_b200.lb_base import LBSim, LBForcedSim, ScalarField, VectorField  # noqa: F401


class LBFluidSim(LBSim):
    @classmethod
    def add_options(cls, group, dim):
        """lb_single.py:21-50."""
        group.add_argument('--visc', type=float, default=1.0, help='numerical viscosity')
        group.add_argument('--model', help='LB collision model to use', type=str,
                           choices=['bgk', 'mrt'], default='bgk')

    @classmethod
    def fields(cls):
        return [ScalarField('rho'), VectorField('v')]

    def update_context(self, ctx):
        """lb_single.py:52-70.  tau = (6 visc + 1) / 2 (sym.relaxation_time); the kernels take the viscosity
        and form the relaxation rates themselves, as the generated code does."""
        super(LBFluidSim, self).update_context(ctx)
        cfg = self.config
        ctx.update(simtype='lbm', model=cfg.model, visc=cfg.visc, tau=(6.0 * cfg.visc + 1.0) / 2.0)
        for key, default in (('subgrid', 'none'), ('regularized', False), ('incompressible', False)):
            ctx[key] = getattr(cfg, key, default)
