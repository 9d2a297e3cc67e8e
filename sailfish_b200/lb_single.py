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
