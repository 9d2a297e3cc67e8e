"""Runs a simulation on a block decomposition through the native controller and assembles the global fields
(TEST INFRASTRUCTURE shared by the CPU stand-in tests and the GPU tests)."""
import numpy as np


def run_blocks(cls, geo, args, steps, replay=False, local=False):
    from sailfish_b200.controller import LBSimulationController
    ctrl = LBSimulationController(cls, geo)
    ctrl.setup(args + ['--quiet'], local=local)
    if replay:
        ctrl.run_steps(steps, save_last=True)
    else:
        for k in range(steps):
            ctrl.step_all(k == steps - 1)
    ctrl.sync()
    return ctrl, collect(ctrl)


def collect(ctrl):
    """[(location, size, [rho, vx, vy(, vz)] over the real nodes)] of the local subdomains."""
    out = []
    for r in ctrl.runners:
        r._fields_to_host()
        sl = r._spec._nonghost_slice
        f = [np.array(r.field_base(r._sim.rho))[sl]] + [np.array(r.field_base(c))[sl] for c in r._sim.v]
        out.append((tuple(r._spec.location), tuple(r._spec.size), np.array(f)))
    return out


def merge(pieces, gsize):
    """Global [field, z, y, x] array from the pieces of all subdomains."""
    nf = pieces[0][2].shape[0]
    out = np.full((nf,) + tuple(reversed(gsize)), np.nan, dtype=np.float32)
    for loc, size, f in pieces:
        sl = tuple(slice(l, l + s) for l, s in zip(reversed(loc), reversed(size)))
        out[(slice(None),) + sl] = f
    return out
