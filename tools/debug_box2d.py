import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import util
from oracle import oracle
from sailfish_b200.runtime import DeviceLattice
np.set_printoptions(linewidth=250, precision=4)
for case in ('box2d_bgk_AB',):
    c = util.load_case(case)
    p = util.params_from_case(c)
    dev = DeviceLattice(p, c['map'], c['rho0'], c['v0'])
    sim = oracle.Simulation(p, c['map'], c['rho0'], c['v0'], dtype=np.float32)
    # collide only
    dev.collide(); 
    oracle.collide_and_propagate(sim.p, sim.map, sim.dist_a, sim.dist_b, sim.rho, sim.v, 0, 0)
    dev.iteration = 1
    g = dev.get_dist(); o = sim.dist_b.reshape(g.shape)
    d = np.abs(g - o); d[~np.isfinite(d)] = 0
    print('after collide: max', d.max(), np.argwhere(d > 1e-5)[:10].tolist())
    nf = np.isfinite(g) != np.isfinite(o)
    print('finite mismatch', np.argwhere(nf)[:20].tolist())
    dev.iteration = 0
    for axis in (0, 1):
        dev.lib.call('slf_apply_periodic_boundary_conditions', dev.plan, int(dev.d_dist_b), axis, 0, dev.stream.handle)
        oracle.apply_pbc(sim.p, sim.dist_b, axis, False)
        dev.iteration = 1
        g = dev.get_dist(); o = sim.dist_b.reshape(g.shape)
        dev.iteration = 0
        d = np.abs(g - o); d[~np.isfinite(d)] = 0
        print('after pbc axis', axis, 'max', d.max(), np.argwhere(d > 1e-5)[:16].tolist())
        print('gpu slot1 col x=1:', g[1, :, 1]); print('orc slot1 col x=1:', o[1, :, 1]); print('gpu slot1 col x=13:', g[1, :, 13])
