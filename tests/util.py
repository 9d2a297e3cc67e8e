"""Shared helpers for the test-suite (fixtures -> SlfParams, oracle runs, norms)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')

from sailfish_b200 import params as P  # noqa: E402


def load_case(name):
    return np.load(os.path.join(GOLD, 'steps_%s.npz' % name))


def params_from_case(c, access_pattern=None):
    type_ids = {str(n): int(i) for n, i in zip(c['type_names'], c['type_ids'])}
    accel = [float(a) for a in c['accel']]
    shan_chen = None
    if 'simtype' in c.files and str(c['simtype']) == 'shan-chen':
        g = [float(x) for x in c['sc_g']]
        shan_chen = dict(tau_phi=float(c['tau_phi']), G11=g[0], G12=g[1], G21=g[2], G22=g[3],
                         potential=str(c['sc_potential']),
                         accel1=[float(a) for a in c['accel1']] if 'accel1' in c.files else None)
    return P.make_params(
        str(c['grid']), str(c['model']), access_pattern or str(c['access_pattern']),
        [int(x) for x in c['lat']], [int(x) for x in c['arr']], [int(x) for x in c['periodic']],
        [int(x) for x in c['bits']], type_ids, [float(x) for x in c['node_params']],
        float(c['visc']), accel, use_link_tags=bool(c['use_link_tags']), shan_chen=shan_chen)


def real_mask(c):
    """Boolean mask (array shape) of the non-ghost, non-padding nodes."""
    lat = [int(x) for x in c['lat']]
    shape = c['map'].shape
    m = np.zeros(shape, bool)
    if len(shape) == 3:
        m[1:lat[2] - 1, 1:lat[1] - 1, 1:lat[0] - 1] = True
    else:
        m[1:lat[1] - 1, 1:lat[0] - 1] = True
    return m


def rel_linf(a, b):
    """||a-b||_inf / ||b||_inf over finite entries of b (the parity metric of
    SURVEY.md section 7 'Hard parts')."""
    ok = np.isfinite(b)
    d = np.abs(a[ok] - b[ok])
    if d.size == 0:
        return 0.0
    scale = np.max(np.abs(b[ok]))
    return float(np.max(d) / scale) if scale > 0 else float(np.max(d))


SC_CASES = ['sc3d', 'sc3d_walls', 'sc2d_classic', 'sc2d_gravity', 'sc3d_gravity']
CASES = ['ldc2d_bgk', 'ldc2d_mrt', 'ldc3d_bgk', 'ldc3d_mrt', 'ldc3d_q27', 'poi3d_press_mrt',
         'poi3d_force_bgk', 'poi3d_force_mrt', 'box3d_bgk', 'box3d_q27', 'box2d_bgk']

# step goldens on tests/geometry_zoo.py (half-way bounce-back, equilibrium velocity, regularised density,
# do-nothing nodes, obstacles with unused cores): name_AP where tools/gen_golden_steps.py produced one.  The
# duct (half-way walls on the outermost layer, no unused layer behind them -- the configuration the reference's
# FIXME in binary_shan_chen.mako / boundary.mako warns about) is pinned in AB only; its AA run is covered by the
# AA == AB invariant.
ZOO_STEP_CASES = sorted(f[len('steps_'):-len('.npz')] for f in os.listdir(os.path.join(ROOT, 'tests', 'golden'))
                        if f.startswith('steps_zoo_'))

