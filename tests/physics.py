"""Physics regression cases of the north star (regtest/ldc_3d.py, regtest/poiseuille.py of the reference),
written against the native runner so that they run on the GPU (tests/test_gpu_physics.py) and, at reduced
size, on the CPU stand-in device (tests/test_physics_fake.py).  TEST INFRASTRUCTURE."""
import os

import numpy as np

import util
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.examples import ldc_3d, poiseuille_3d

GOLD = np.load(os.path.join(util.GOLD, 'physics.npz'))


# ---- lid-driven cavity, Re = 400 (regtest/ldc_3d.py:22-66) ---------------------------------------
def ldc3d_re400(n, steps, access_pattern='AA', model='bgk'):
    """Runs the cubic cavity at Re = 400 (visc = (n - 2) max_v / Re, regtest/ldc_3d.py:37-40) and returns the
    two centre-line profiles of regtest/ldc_3d.py:53-66 plus their largest deviation from the points digitised
    from Sheu & Tsai (regtest/ldc_golden/re400_{horiz,vert}; plotted against the result at :84-88)."""
    max_v = ldc_3d.Cavity.lid_speed
    visc = (n - 2) * max_v / 400.0
    ctrl = LBSimulationController(ldc_3d.LDCSim)
    ctrl.setup(['--lat_nx=%d' % n, '--lat_ny=%d' % n, '--lat_nz=%d' % n, '--visc=%r' % visc,
                '--access_pattern=' + access_pattern, '--model=' + model, '--grid=D3Q19', '--quiet'])
    r = ctrl.runners[0]
    r.run_steps(steps, save_last=True)
    ctrl.sync()
    r._fields_to_host()
    vx, vz = np.array(r._sim.vx), np.array(r._sim.vz)
    rho = np.array(r._sim.rho)
    h = n // 2
    res_vx = (vx[:, h, h] + vx[:, h - 1, h - 1]) / 2 / max_v          # along z through the centre
    res_vz = (vz[h, h, :] + vz[h - 1, h - 1, :]) / 2 / max_v          # along x through the centre
    pos = np.linspace(-1.0, 1.0, n)
    horiz, vert = GOLD['ldc_re400_horiz'], GOLD['ldc_re400_vert']
    g_vx, g_z = 2 * (horiz[:, 0] - 0.5), -2 * (horiz[:, 1] - 0.5)
    g_x, g_vz = 2 * (vert[:, 0] - 0.5), -2 * (vert[:, 1] - 0.5)
    dev_vx = float(np.max(np.abs(np.interp(g_z, pos, res_vx) - g_vx)))
    dev_vz = float(np.max(np.abs(np.interp(g_x, pos, res_vz) - g_vz)))
    wet = r._subdomain.fluid_map()
    return dict(res_vx=res_vx, res_vz=res_vz, dev_vx=dev_vx, dev_vz=dev_vz, visc=visc,
                finite=bool(np.isfinite(vx[wet]).all() and np.isfinite(rho[wet]).all()),
                vx_min=float(res_vx.min()), vz_min=float(res_vz.min()), vz_max=float(res_vz.max()))


# ---- Poiseuille flow in a pipe, force driven, full-way bounce-back (regtest/poiseuille.py:38-62;
#      table regtest/results/poiseuille3d/D3Q19/{bgk,mrt}/force/single/fullbb.dat) ---------------------------
class StationaryPipe(poiseuille_3d.Pipe):
    """examples/poiseuille_3d.py:66-96 with --stationary --drive=force: start from the analytic profile."""

    def _radius_sq(self, hx, hy, hz):
        coords = dict(x=(hx, self.gx), y=(hy, self.gy), z=(hz, self.gz))
        (a, na), (b, nb) = [coords[ax] for ax in 'xyz' if ax != self.config.flow_direction]
        return (a - na / 2.0 + 0.5) ** 2 + (b - nb / 2.0 + 0.5) ** 2

    @classmethod
    def velocity_profile(cls, config, rsq):
        w = cls.channel_width(config)
        return cls.max_v / (w / 2.0) ** 2 * ((w / 2.0) ** 2 - rsq)

    def initial_conditions(self, sim, hx, hy, hz):
        sim.rho[:] = 1.0
        rsq = np.broadcast_to(self._radius_sq(hx, hy, hz), sim.rho.shape)
        inside = rsq <= (self.channel_width(self.config) / 2.0) ** 2
        comp = dict(x=sim.vx, y=sim.vy, z=sim.vz)[self.config.flow_direction]
        comp[inside] = self.velocity_profile(self.config, rsq[inside])


class StationaryPoiseuilleSim(poiseuille_3d.PoiseuilleSim):
    subdomain = StationaryPipe


def poiseuille3d_max_velocity_error(visc, model, width=64, length=8, steps=None, access_pattern='AA', direction='x'):
    """max(v_sim) / max(v_theory) - 1 after int(100 / visc) steps from the stationary profile
    (regtest/poiseuille.py:48-75)."""
    steps = steps or int(100 / visc)
    size = dict(x=width, y=width, z=width)
    size[direction] = length
    ctrl = LBSimulationController(StationaryPoiseuilleSim)
    ctrl.setup(['--lat_nx=%d' % size['x'], '--lat_ny=%d' % size['y'], '--lat_nz=%d' % size['z'], '--visc=%r' % visc,
                '--model=' + model, '--drive=force', '--flow_direction=' + direction, '--stationary',
                '--access_pattern=' + access_pattern, '--quiet'])
    r = ctrl.runners[0]
    r.run_steps(steps, save_last=True)
    ctrl.sync()
    r._fields_to_host()
    v = np.array(dict(x=r._sim.vx, y=r._sim.vy, z=r._sim.vz)[direction])
    wet = r._subdomain.fluid_map()
    sd = r._subdomain
    rsq = np.broadcast_to(sd._radius_sq(*sd._get_mgrid()), v.shape)
    th = StationaryPipe.velocity_profile(ctrl.config, rsq)
    err = float(np.max(v[wet]) / np.max(th[wet]) - 1.0)
    return dict(err=err, steps=steps, finite=bool(np.isfinite(v[wet]).all()),
                profile_err=float(np.max(np.abs(v[wet] - th[wet])) / np.max(th[wet])))


def reference_poiseuille_error(visc, model):
    """The reference's own result at this viscosity (log-interpolated from its 30-point table)."""
    t = GOLD['poiseuille3d_%s_force_fullbb' % model]
    return float(np.interp(np.log(visc), np.log(t[:, 0]), t[:, 1]))
