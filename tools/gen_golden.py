"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference
(/root/reference, imported with the py3.12 shims of sailfish_b200/compat.py).

Run in the build container only:  python tools/gen_golden.py [what ...]
   what in {lattices, formulas, geometry, steps}; default: all.

Outputs (all small .npz files, committed):
  lattices.npz  -- basis / weights / opposite / names / MRT matrices straight from
                   sailfish.sym (sym.py:61-428, :944-1045) + interblock/prop sets.
  formulas.npz  -- every closed-form expression of the hot path evaluated by
                   lambdifying the reference's own SymPy objects on seeded random
                   inputs: f^eq (sym_equilibrium.py:90-120), velocity/density BC
                   density+velocity (sym.py:573-682), non-equilibrium bounce-back
                   (sym.py:750-766), regularisation (sym.py:882-891), Guo force
                   (sym_force.py:121-160), MRT equilibria/rates (sym.py:98-149,361-406).
  geometry_*.npz-- encoded node maps, node_params, bit layout and initial fields
                   produced by the reference's Subdomain/GeoEncoderConst for small
                   versions of the BASELINE configs.
  steps_*.npz   -- N-step results of tools/ref_stepper.py (NumPy restatement of the
                   kernel control flow whose formulas are the reference's SymPy
                   expressions) for those geometries.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness  # noqa: E402  (installs shims, puts reference on sys.path)

import numpy as np  # noqa: E402
import sympy  # noqa: E402
from sailfish import sym, sym_equilibrium, sym_force  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
GRIDS = {'D2Q9': sym.D2Q9, 'D3Q19': sym.D3Q19, 'D3Q27': sym.D3Q27}


class _Cfg(object):
    incompressible = False
    minimize_roundoff = False


def gen_lattices():
    out = {}
    for name, g in GRIDS.items():
        out[name + '_basis'] = np.array([[int(c) for c in e] for e in g.basis], dtype=np.int32)
        out[name + '_weights_num'] = np.array([int(w.p) for w in g.weights], dtype=np.int64)
        out[name + '_weights_den'] = np.array([int(w.q) for w in g.weights], dtype=np.int64)
        out[name + '_opposite'] = np.array(g.idx_opposite, dtype=np.int32)
        out[name + '_names'] = np.array(g.idx_name)
        out[name + '_dir2vecidx'] = np.array([g.dir2vecidx[d] for d in range(1, 2 * g.dim + 1)], dtype=np.int32)
        out[name + '_bb_swap'] = np.array(sorted(sym.bb_swap_pairs(g)), dtype=np.int32)
        for ax in range(g.dim):
            for d in (-1, 1):
                out['%s_prop_%d_%d' % (name, ax, d)] = np.array(sym.get_prop_dists(g, d, ax), dtype=np.int32)
                vec = [0] * g.dim
                vec[ax] = d
                out['%s_interblock_%d_%d' % (name, ax, d)] = np.array(
                    sym.get_interblock_dists(g, vec), dtype=np.int32)
        for o in range(1, 2 * g.dim + 1):
            out['%s_missing_%d' % (name, o)] = np.array(sym.get_missing_dists(g, o), dtype=np.int32)
        if hasattr(g, 'mrt_matrix'):
            out[name + '_mrt_matrix'] = np.array(g.mrt_matrix.tolist(), dtype=np.int64)
            out[name + '_mrt_names'] = np.array(g.mrt_names)
    np.savez_compressed(os.path.join(GOLD, 'lattices.npz'), **out)
    print('lattices.npz written')


def _lam(args, expr):
    return sympy.lambdify(args, expr, modules='numpy')


def gen_formulas(n=64, seed=20260924):
    rng = np.random.RandomState(seed)
    S = sym.S
    cfg = _Cfg()
    out = {}
    for name, g in GRIDS.items():
        dim, Q = g.dim, g.Q
        rho = 1.0 + 0.1 * (rng.rand(n) - 0.5)
        v = 0.1 * (rng.rand(dim, n) - 0.5)
        f = np.array([float(w) for w in g.weights])[:, None] * (1.0 + 0.2 * (rng.rand(Q, n) - 0.5))
        a = 1e-3 * (rng.rand(dim, n) - 0.5)
        out[name + '_rho'], out[name + '_v'], out[name + '_f'], out[name + '_a'] = rho, v, f, a
        vs = [S.vx, S.vy, S.vz][:dim]
        eq = sym_equilibrium.bgk_equilibrium(g, cfg).expression
        feq = np.array([_lam([S.rho] + vs, e)(rho, *v) * np.ones(n) for e in eq])
        out[name + '_feq'] = feq

        fsyms = [sympy.Symbol('fi->%s' % x) for x in g.idx_name]
        # Velocity / density BC macro (boundary.mako:425-459).
        for o in range(1, 2 * dim + 1):
            ex = sym.ex_rho(g, 'fi', False, missing_dir=o)
            out['%s_vbc_rho_%d' % (name, o)] = _lam([S.rho] + vs, ex)(rho, *v)
            dv = []
            for d in range(dim):
                ex = sym.ex_velocity(g, 'fi', d, cfg, missing_dir=o, par_rho='par_rho')
                dv.append(_lam([S.rho, sympy.Symbol('par_rho')], ex)(rho, 1.0 + 0.05 * np.ones(n)) * np.ones(n))
            out['%s_dbc_v_%d' % (name, o)] = np.array(dv)
            # Non-equilibrium bounce-back (sym.py:750-766).
            res = f.copy()
            for arg, val in sym.noneq_bb(g, o, eq):
                i = fsyms.index(arg)
                res[i] = _lam(fsyms + [S.rho] + vs, val)(*f, rho, *v)
            out['%s_noneq_bb_%d' % (name, o)] = res
        # Second moments (boundary.mako:289-304).
        comps = [(x, y) for x in range(dim) for y in range(x, dim)]
        flux = np.array([_lam(fsyms, sym.ex_flux(g, 'fi', x, y, cfg))(*f) * np.ones(n) for x, y in comps])
        eqflux = np.array([_lam([S.rho] + vs, sym.ex_eq_flux(g, x, y))(rho, *v) for x, y in comps])
        out[name + '_flux'] = flux
        out[name + '_eqflux'] = eqflux
        fl = [sympy.Symbol('flux[%d]' % i) for i in range(len(comps))]
        noneq = flux - eqflux
        out[name + '_reg'] = np.array([_lam(fl, e)(*noneq) * np.ones(n) for e in sym.reglb_flux_tensor(g)])
        # Guo forcing (sym_force.py:121-160); pref evaluated with tau0 symbol.
        tau = 0.8
        ea = [getattr(S, 'g0ea' + c) for c in 'xyz'][:dim]
        pref = sympy.Symbol('pref')
        pref_ex = sym_force.guo_external_force_pref(g, cfg, grid_num=0)
        pref_v = _lam([S.rho, S.tau0], pref_ex)(rho, tau)
        out[name + '_guo_pref'] = pref_v
        out[name + '_guo'] = np.array([
            _lam([pref] + vs + ea, e)(pref_v, *v, *a) * np.ones(n)
            for e in sym_force.guo_external_force(g, grid_num=0)])
        out[name + '_guo_tau'] = np.float64(tau)
        # MRT (sym.py:716-735, relaxation_mrt.mako:31-97).
        if hasattr(g, 'mrt_matrix'):
            visc = 0.05
            m = np.array(g.mrt_matrix.tolist(), dtype=np.float64) @ f
            out[name + '_mrt_m'] = m
            minv = np.array(g.mrt_matrix.inv().tolist(), dtype=np.float64)
            out[name + '_mrt_inv'] = minv
            loc = {}
            for eqn in g.mrt_eq_symbols:
                loc[eqn.lhs] = eqn.rhs.subs(S.rho0, S.rho).subs(S.visc, visc)
            msyms = [g.mx, g.my, g.mz][:dim]
            mom = [m[g.mrt_names.index(k)] for k in ('mx', 'my', 'mz')[:dim]]
            meq, rates = [], []
            for i, e in enumerate(g.mrt_equilibrium):
                e = sympy.sympify(e).subs(loc)
                meq.append(_lam([S.rho] + msyms, e)(m[0], *mom) * np.ones(n))
                r = sympy.sympify(g.mrt_collision[i]).subs(loc)
                rates.append(float(r.subs(S.visc, visc)))
            out[name + '_mrt_eq'] = np.array(meq)
            out[name + '_mrt_rates'] = np.array(rates)
            out[name + '_mrt_visc'] = np.float64(visc)
    np.savez_compressed(os.path.join(GOLD, 'formulas.npz'), **out)
    print('formulas.npz written')


if __name__ == '__main__':
    what = sys.argv[1:] or ['lattices', 'formulas', 'geometry', 'steps']
    os.makedirs(GOLD, exist_ok=True)
    if 'lattices' in what:
        gen_lattices()
    if 'formulas' in what:
        gen_formulas()
    if 'geometry' in what or 'steps' in what:
        import gen_golden_steps
        gen_golden_steps.main(what)
