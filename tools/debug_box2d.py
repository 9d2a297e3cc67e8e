import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import util
from oracle import oracle
from sailfish_b200.runtime import DeviceLattice
for case in ('box2d_bgk_AB', 'box2d_bgk_AA'):
    c = util.load_case(case)
    p = util.params_from_case(c)
    dev = DeviceLattice(p, c['map'], c['rho0'], c['v0'])
    sim = oracle.Simulation(p, c['map'], c['rho0'], c['v0'], dtype=np.float32)
    for step in range(3):
        dev.step(); sim.step()
        g = dev.get_dist(); o = sim.current_dist().reshape(g.shape)
        d = np.abs(g - o); d[~np.isfinite(d)] = 0
        bad = np.argwhere(d > 1e-5)
        print(case, 'step', step, 'max', d.max(), 'nbad', len(bad), 'first', bad[:12].tolist())
