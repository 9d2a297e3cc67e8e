"""Binary-fluid model classes: mirror of sailfish/lb_binary.py:14-137 (LBBinaryFluidBase) and
:375-517 (LBBinaryFluidShanChen) -- BASELINE config 5."""
from sailfish_b200.lb_base import LBSim, LBForcedSim, ScalarField, VectorField
from sailfish_b200.subdomain_runner import NNSubdomainRunner


class LBBinaryFluidBase(LBSim):
    subdomain_runner = NNSubdomainRunner
    nonlocality = 1

    def __init__(self, config):
        super(LBBinaryFluidBase, self).__init__(config)
        self.grids.append(self.grid)

    @classmethod
    def add_options(cls, group, dim):
        group.add_argument('--tau_phi', type=float, default=1.0, help='relaxation time for the phase field')


class LBBinaryFluidShanChen(LBBinaryFluidBase, LBForcedSim):
    """lb_binary.py:375-420: two BGK lattices coupled by the Shan-Chen force (G11, G12 = G21, G22)."""

    @classmethod
    def fields(cls):
        return [ScalarField('rho'), ScalarField('phi'), VectorField('v')]

    @classmethod
    def add_options(cls, group, dim):
        group.add_argument('--visc', type=float, default=1.0, help='numerical viscosity')
        group.add_argument('--G11', type=float, default=0.0)
        group.add_argument('--G12', type=float, default=0.0)
        group.add_argument('--G22', type=float, default=0.0)
        group.add_argument('--sc_potential', type=str, choices=['linear', 'classic'], default='linear')
        group.add_argument('--model', type=str, choices=['bgk'], default='bgk')
