import sys, numpy as np, ctypes
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import util
from sailfish_b200 import backend_b200 as bk
bk.LIB_PATH = bk.LIB_PATH.replace('libslf_b200.so', 'libslf_b200_dbg.so')
bk._LIB = bk.Lib(bk.LIB_PATH)
from oracle import oracle
from sailfish_b200.runtime import DeviceLattice
c = util.load_case('box2d_bgk_AA'); p = util.params_from_case(c)
dev = DeviceLattice(p, c['map'], c['rho0'], c['v0'])
shape = (p.Q,) + c['map'].shape
d0 = np.random.RandomState(7).rand(*shape).astype(np.float32)
for swap in (0, 1):
    dev.set_dist(d0)
    dev.lib.call('slf_apply_periodic_boundary_conditions', dev.plan, int(dev.d_dist_a), 0, swap, dev.stream.handle)
    g = dev.get_dist()
    o = d0.copy().reshape(-1); oracle.apply_pbc(p, o, 0, bool(swap)); o = o.reshape(shape)
    print('swap', swap, 'mismatch', np.argwhere(g != o)[:30].tolist())
