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

    def update_context(self, ctx):
        """lb_binary.py:130-133."""
        super(LBBinaryFluidBase, self).update_context(ctx)
        ctx.update(tau_phi=self.config.tau_phi, model='bgk')


class LBBinaryFluidShanChen(LBBinaryFluidBase, LBForcedSim):
    """lb_binary.py:375-420: two BGK lattices coupled by the Shan-Chen force (G11, G12 = G21, G22)."""

    def __init__(self, config):
        """lb_binary.py:382-387."""
        super(LBBinaryFluidShanChen, self).__init__(config)
        for a in (0, 1):
            for b in (0, 1):
                self.add_force_coupling(a, b, 'G%d%d' % (a + 1, b + 1))

    def constants(self):
        """lb_binary.py:389-395: the cross coupling is symmetric (G21 = G12)."""
        ret = super(LBBinaryFluidShanChen, self).constants()
        cfg = self.config
        ret.update(G11=cfg.G11, G12=cfg.G12, G21=cfg.G12, G22=cfg.G22)
        return ret

    @classmethod
    def fields(cls):
        return [ScalarField('rho'), ScalarField('phi'), VectorField('v')]

    def update_context(self, ctx):
        """lb_binary.py:416-421."""
        super(LBBinaryFluidShanChen, self).update_context(ctx)
        cfg = self.config
        ctx.update(simtype='shan-chen', sc_potential=cfg.sc_potential, visc=cfg.visc,
                   tau=(6.0 * cfg.visc + 1.0) / 2.0)

    @classmethod
    def add_options(cls, group, dim):
        group.add_argument('--visc', type=float, default=1.0, help='numerical viscosity')
        group.add_argument('--G11', type=float, default=0.0)
        group.add_argument('--G12', type=float, default=0.0)
        group.add_argument('--G22', type=float, default=0.0)
        group.add_argument('--sc_potential', type=str, choices=['linear', 'classic'], default='linear')
        group.add_argument('--model', type=str, choices=['bgk'], default='bgk')
