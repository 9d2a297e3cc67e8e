"""NumPy stepper whose ARITHMETIC is the reference's own SymPy objects.

Purpose: produce whole-step golden vectors without Mako/pycuda/GPU.  The control
flow follows templates/models/lb_single_fluid.mako:161-229 and the helpers it
inlines; every formula (moments, equilibria, boundary closures, MRT transform,
Guo forcing) is obtained by ``sympy.lambdify`` from sailfish/sym.py,
sym_equilibrium.py and sym_force.py -- the very expressions the reference prints
into its CUDA source.  Geometry (node map, node_params, bit layout, type flags)
comes from the reference's Subdomain/GeoEncoderConst/node_type classes.

Runs in float64; build-container only (needs /root/reference).
"""
import numpy as np
import sympy

from sailfish import sym, sym_equilibrium, sym_force
import sailfish.node_type as nt


def _lam(args, expr):
    return sympy.lambdify(args, expr, modules='numpy')


class RefStepper(object):
    def __init__(self, runner, ctx, access_pattern=None):
        self.ctx = ctx
        self.grid = g = ctx['grid']
        self.cfg = ctx['config']
        self.aa = (access_pattern or ctx['access_pattern']) == 'AA'
        self.Q, self.D = g.Q, g.dim
        self.shape = tuple(runner._physical_size)            # ([nz], ny, arr_nx)
        self.lat = tuple(runner._lat_size)
        self.n = int(np.prod(self.shape))
        S = sym.S
        self.vs = [S.vx, S.vy, S.vz][:self.D]
        self.fs = [sympy.Symbol('fi->%s' % x) for x in g.idx_name]
        self.map = runner._subdomain.encoded_map().astype(np.uint32).reshape(self.shape).copy()
        remap = ctx['type_id_remap']
        self.cls_of = {int(remap[c.id]): c for c in ctx['node_types']}
        tmask = ctx['nt_type_mask']
        self.type = self.map & tmask
        self.orient = self.map >> (ctx['nt_misc_shift'] + ctx['nt_param_shift'] + ctx['nt_scratch_shift'])
        self.pidx = (self.map >> ctx['nt_misc_shift']) & ((1 << ctx['nt_param_shift']) - 1)
        self.node_params = np.array([float(x) for x in ctx['node_params']] + [0.0] * 3)
        self.periodic = list(ctx['block_periodicity'])[:self.D]

        def flag(name):
            out = np.zeros(self.shape, bool)
            for tid, c in self.cls_of.items():
                if getattr(c, name):
                    out |= (self.type == tid)
            return out
        self.excluded = flag('excluded')
        self.prop_only = flag('propagation_only')
        self.wet = flag('wet_node')
        self.std_macro = flag('standard_macro')
        x_ok = np.zeros(self.shape, bool)
        x_ok[..., :self.lat[-1]] = True
        self.active = (~self.excluded) & (~self.prop_only) & x_ok

        def is_cls(c):
            out = np.zeros(self.shape, bool)
            for tid, k in self.cls_of.items():
                if k is c:
                    out |= (self.type == tid)
            return out
        self.is_cls = is_cls

        # --- lambdified reference expressions -------------------------------------
        eq = sym_equilibrium.bgk_equilibrium(g, self.cfg).expression
        self.eq_exprs = eq
        self.feq = [_lam([S.rho] + self.vs, e) for e in eq]
        self.rho_std = _lam(self.fs, sym.ex_rho(g, 'fi', False))
        self.mom = [_lam(self.fs, sym.ex_velocity(g, 'fi', d, self.cfg, momentum=True)) for d in range(self.D)]
        self.rho_vbc = {o: _lam([S.rho] + self.vs, sym.ex_rho(g, 'fi', False, missing_dir=o))
                        for o in range(1, 2 * self.D + 1)}
        par = sympy.Symbol('par_rho')
        self.v_dbc = {o: [_lam([S.rho, par], sym.ex_velocity(g, 'fi', d, self.cfg, missing_dir=o, par_rho='par_rho'))
                          for d in range(self.D)] for o in range(1, 2 * self.D + 1)}
        self.missing = {o: list(sym.get_missing_dists(g, o)) for o in range(1, 2 * self.D + 1)}
        self.noneq = {}
        for o in range(1, 2 * self.D + 1):
            self.noneq[o] = [(self.fs.index(a), _lam(self.fs + [S.rho] + self.vs, val))
                             for a, val in sym.noneq_bb(g, o, eq)]
        # Zou-He closures (boundary.mako:321-382, sym.py:770-815): momentum fix-up after the non-equilibrium bounce-back
        nv = [sympy.Symbol(n) for n in ('nvx', 'nvy', 'nvz')]
        self.zh_fix = {}
        for o in range(1, 2 * self.D + 1):
            steps = []
            for lhs, rhs in sym.zouhe_fixup(g, o):
                lhs = sympy.Symbol(str(lhs))
                tgt = ('nv', nv.index(lhs)) if lhs in nv else ('f', self.fs.index(lhs))
                steps.append((tgt, _lam(self.fs + nv, rhs)))
            self.zh_fix[o] = steps
        comps = [(a, b) for a in range(self.D) for b in range(a, self.D)]
        self.flux = [_lam(self.fs, sym.ex_flux(g, 'fi', a, b, self.cfg)) for a, b in comps]
        self.eqflux = [_lam([S.rho] + self.vs, sym.ex_eq_flux(g, a, b)) for a, b in comps]
        fl = [sympy.Symbol('flux[%d]' % i) for i in range(len(comps))]
        self.reg = [_lam(fl, e) for e in sym.reglb_flux_tensor(g)]
        self.tau = float(ctx['tau'])
        self.model = ctx['model']
        # forces
        self.accel = None
        forces = ctx.get('forces')
        if forces is not None and getattr(forces, 'numeric', None):
            self.accel = np.array(forces.numeric[0][True], dtype=np.float64)
            ea = [getattr(S, 'g0ea' + c) for c in 'xyz'][:self.D]
            pref = sympy.Symbol('pref')
            self.guo_pref = _lam([S.rho, S.tau0], sym_force.guo_external_force_pref(g, self.cfg, 0))
            self.guo = [_lam([pref] + self.vs + ea, e) for e in sym_force.guo_external_force(g, 0)]
        if self.model == 'mrt':
            self.M = np.array(g.mrt_matrix.tolist(), dtype=np.float64)
            self.Minv = np.array(g.mrt_matrix.inv().tolist(), dtype=np.float64)
            loc = {}
            for eqn in g.mrt_eq_symbols:
                loc[eqn.lhs] = eqn.rhs.subs(S.rho0, S.rho).subs(S.visc, float(ctx['visc']))
            msyms = [g.mx, g.my, g.mz][:self.D]
            self.m_idx = [g.mrt_names.index(k) for k in ('mx', 'my', 'mz')[:self.D]]
            self.meq, self.rates = [], []
            for i, e in enumerate(g.mrt_equilibrium):
                r = sympy.sympify(g.mrt_collision[i]).subs(loc)
                self.rates.append(float(r))
                self.meq.append(_lam([S.rho] + msyms, sympy.sympify(e).subs(loc)))
        # --- state ---------------------------------------------------------------
        self.dist_a = np.zeros((self.Q,) + self.shape)
        self.dist_b = None if self.aa else np.zeros((self.Q,) + self.shape)
        self.iteration = 0
        self.rho = None
        self.v = None

    # ------------------------------------------------------------------------------
    def set_initial_conditions(self, rho, v):
        """SetInitialConditions, lb_single_fluid.mako:101-127 (all nodes, no type test)."""
        with np.errstate(all='ignore'):
            for i in range(self.Q):
                val = self.feq[i](rho, *v) * np.ones(self.shape)
                val[..., self.lat[-1]:] = 0.0   # padding is never touched (gx > lat_nx-1 returns)
                self.dist_a[i] = val
                if self.dist_b is not None:
                    self.dist_b[i] = val
        self.rho = np.array(rho, dtype=np.float64).copy()
        self.v = [np.array(c, dtype=np.float64).copy() for c in v]

    def current_dist(self):
        if self.dist_b is None or (self.iteration & 1) == 0:
            return self.dist_a
        return self.dist_b

    def step(self, save_macro=False):
        g, Q, D = self.grid, self.Q, self.D
        it = self.iteration
        odd = self.aa and (it & 1)
        if self.aa:
            din = dout = self.dist_a
        elif (it & 1) == 0:
            din, dout = self.dist_a, self.dist_b
        else:
            din, dout = self.dist_b, self.dist_a
        idx = np.nonzero(self.active)
        N = idx[0].size
        # getDist (geo_helpers.mako:225-276)
        f = np.zeros((Q, N))
        for i in range(Q):
            if not odd:
                f[i] = din[i][idx]
            else:
                e = [int(c) for c in g.basis[i]]
                src = tuple(idx[k] - e[D - 1 - k] for k in range(D))
                f[i] = din[g.idx_opposite[i]][src]
        typ = self.type[idx]
        ori = self.orient[idx].astype(np.int64)
        pid = self.pidx[idx].astype(np.int64)
        std = self.std_macro[idx] | (ori == 0)
        wet = self.wet[idx]

        def m(c):
            return self.is_cls(c)[idx]
        # fixMissingDistributions (boundary.mako:509-630): copy / extrapolation outflow nodes take the populations
        # that would come from outside from their inward neighbours' dist_in (AB only, :626-628)
        for c, two in ((nt.NTCopy, False), (nt.NTYuOutflow, True)):
            if c not in self.ctx['node_types']:
                continue
            assert not self.aa, 'NTCopy / NTYuOutflow are not supported with the AA access pattern'
            sc = m(c)
            for o in range(1, 2 * D + 1):
                s = sc & (ori == o)
                if not s.any():
                    continue
                nvec = [int(x) for x in g.dir_to_vec(o)]
                nb1 = tuple(idx[k][s] + nvec[D - 1 - k] for k in range(D))
                nb2 = tuple(idx[k][s] + 2 * nvec[D - 1 - k] for k in range(D))
                for i in self.missing[o]:
                    f[i, s] = 2.0 * din[i][nb1] - din[i][nb2] if two else din[i][nb1]
        # NTWallTMS (boundary.mako:631-648): target moments from the populations as loaded (the slots of the links
        # that point into the wall hold the bounced-back values written at the end of the previous step), then
        # those populations are replaced by the equilibrium at the target moments
        tms = m(nt.NTWallTMS)
        tg_rho = tg_v = None
        if tms.any():
            with np.errstate(all='ignore'):
                tg_rho = np.zeros(N)
                tg_v = np.zeros((D, N))
                tg_rho[tms] = self.rho_std(*f[:, tms])
                for d in range(D):
                    tg_v[d, tms] = self.mom[d](*f[:, tms]) / tg_rho[tms]
                for sel, missing, _ in self._missing_links(ori, tms):
                    f[missing, sel] = self.feq[missing](tg_rho[sel], *tg_v[:, sel])
        zh_v, zh_d = m(nt.NTZouHeVelocity), m(nt.NTZouHeDensity)
        vel_bc = m(nt.NTEquilibriumVelocity) | m(nt.NTRegularizedVelocity) | zh_v
        den_bc = m(nt.NTEquilibriumDensity) | m(nt.NTRegularizedDensity) | zh_d
        reg_bc = m(nt.NTRegularizedVelocity) | m(nt.NTRegularizedDensity)
        eq_bc = m(nt.NTEquilibriumVelocity) | m(nt.NTEquilibriumDensity)
        # getMacro (boundary.mako:465-507)
        rho = np.zeros(N)
        v = np.zeros((D, N))
        with np.errstate(all='ignore'):
            s = std
            rho[s] = self.rho_std(*f[:, s])
            for d in range(D):
                v[d, s] = self.mom[d](*f[:, s]) / rho[s]
            for o in range(1, 2 * D + 1):
                s = (~std) & (vel_bc | den_bc) & (ori == o)
                if not s.any():
                    continue
                for i in self.missing[o]:
                    f[i, s] = f[g.idx_opposite[i], s]
                r0 = self.rho_std(*f[:, s])
                sv = s & vel_bc
                if sv.any():
                    r0v = self.rho_std(*f[:, sv])
                    for d in range(D):
                        v[d, sv] = self.node_params[pid[sv] + d]
                    rho[sv] = self.rho_vbc[o](r0v, *v[:, sv])
                sd = s & den_bc
                if sd.any():
                    r0d = self.rho_std(*f[:, sd])
                    pr = self.node_params[pid[sd]]
                    for d in range(D):
                        v[d, sd] = self.v_dbc[o][d](r0d, pr) * np.ones(pr.shape)
                    rho[sd] = pr
                # NTZouHeDensity (boundary.mako:487-492): fix the populations up, recompute the moments, impose rho
                sz = s & zh_d
                if sz.any():
                    pz = self.node_params[pid[sz]]
                    self._zouhe_bb(f, sz, o, pz, v[:, sz])
                    rz = self.rho_std(*f[:, sz])
                    for d in range(D):
                        v[d, sz] = self.mom[d](*f[:, sz]) / rz
                    rho[sz] = pz
                del r0
            # precollisionBoundaryConditions (boundary.mako:784-878)
            s = m(nt.NTFullBBWall)
            if s.any():
                for i in sym.bb_swap_pairs(g):
                    j = g.idx_opposite[i]
                    t = f[i, s].copy()
                    f[i, s] = f[j, s]
                    f[j, s] = t
            if nt.NTSlip in self.ctx['node_types']:          # boundary.mako:837-856: specular reflection
                ssl = m(nt.NTSlip)
                for o in range(1, 2 * D + 1):
                    s = ssl & (ori == o)
                    if s.any():
                        for j, k in sym.slip_bb_swap_pairs(g, o):
                            t = f[j, s].copy()
                            f[j, s] = f[k, s]
                            f[k, s] = t
            s = eq_bc
            if s.any():
                for i in range(Q):
                    f[i, s] = self.feq[i](rho[s], *v[:, s])
            if zh_v.any():                       # boundary.mako:812-816
                s0 = zh_v & (ori == 0)
                if s0.any():
                    for i in sym.bb_swap_pairs(g):
                        j = g.idx_opposite[i]
                        t = f[i, s0].copy()
                        f[i, s0] = f[j, s0]
                        f[j, s0] = t
                for o in range(1, 2 * D + 1):
                    s = zh_v & (ori == o)
                    if s.any():
                        self._zouhe_bb(f, s, o, rho[s], v[:, s])
            if reg_bc.any():
                s0 = reg_bc & (ori == 0)
                if s0.any():
                    for i in sym.bb_swap_pairs(g):
                        j = g.idx_opposite[i]
                        t = f[i, s0].copy()
                        f[i, s0] = f[j, s0]
                        f[j, s0] = t
                for o in range(1, 2 * D + 1):
                    s = reg_bc & (ori == o)
                    if not s.any():
                        continue
                    fin = f[:, s].copy()
                    for i, fn in self.noneq[o]:
                        f[i, s] = fn(*fin, rho[s], *v[:, s])
                    noneq = [self.flux[k](*f[:, s]) - self.eqflux[k](rho[s], *v[:, s])
                             for k in range(len(self.flux))]
                    for i in range(Q):
                        val = self.feq[i](rho[s], *v[:, s]) + self.reg[i](*noneq)
                        f[i, s] = np.maximum(1e-7, val)
            if self.aa and nt.NTDoNothing in self.ctx['node_types']:
                sdn = m(nt.NTDoNothing)
                for o in range(1, 2 * D + 1):
                    s = sdn & (ori == o)
                    if not s.any():
                        continue
                    sub = tuple(a[s] for a in idx)
                    for i in self.missing[o]:
                        if odd:
                            dout[i][sub] = f[i, s]
                        else:
                            oi = g.idx_opposite[i]
                            e = [int(c) for c in g.basis[oi]]
                            tgt = tuple(sub[k] + e[D - 1 - k] for k in range(D))
                            dout[oi][tgt] = f[i, s]
            # relaxate (relaxation.mako:183-202)
            if self.ctx.get('relaxation_enabled', True):
                s = wet
                if self.model == 'mrt':
                    self._mrt(f, v, s, m)
                else:
                    self._bgk(f, rho, v, s)
            # postcollision: half-way bounce-back (boundary.mako:653-683)
            if nt.NTHalfBBWall in self.ctx['node_types']:
                self._halfbb(f, dout, idx, ori, m(nt.NTHalfBBWall), odd)
            if tms.any():
                # boundary.mako:696-722: add f^eq(target) - f^eq(instantaneous), then store like half-way bounce-back
                for i in range(Q):
                    f[i, tms] += self.feq[i](tg_rho[tms], *tg_v[:, tms])
                for i in range(Q):
                    f[i, tms] -= self.feq[i](rho[tms], *v[:, tms])
                self._halfbb(f, dout, idx, ori, tms, odd)
        # save_macro_fields (kernel_common.mako:213-240)
        if save_macro:
            sub = tuple(a[wet] for a in idx)
            self.rho[sub] = rho[wet]
            for d in range(D):
                self.v[d][sub] = v[d, wet]
        # propagate (propagation.mako:384-421, bounds :36-41)
        if not self.ctx.get('propagation_enabled', True):
            for i in range(Q):
                dout[i][idx] = f[i]
        elif self.aa and not odd:
            for i in range(Q):
                dout[g.idx_opposite[i]][idx] = f[i]
        else:
            for i in range(Q):
                e = [int(c) for c in g.basis[i]]
                ok = np.ones(N, bool)
                for k in range(D):
                    c = e[D - 1 - k]
                    if c > 0:
                        ok &= idx[k] < self.lat[k] - 1
                    elif c < 0:
                        ok &= idx[k] > 0
                tgt = tuple(idx[k][ok] + e[D - 1 - k] for k in range(D))
                dout[i][tgt] = f[i, ok]
        # periodic boundary conditions (subdomain_runner.py:984-1009)
        self._pbc(dout, swap=(self.aa and not odd))
        self.iteration += 1

    # ------------------------------------------------------------------------------
    def _zouhe_bb(self, f, s, o, rho, v):
        """zouhe_bb (boundary.mako:347-381) on the nodes selected by `s` with orientation o: non-equilibrium
        bounce-back of the unknown populations, then the tangential momentum excess is spread over them."""
        D = self.D
        fin = f[:, s].copy()
        for i, fn in self.noneq[o]:
            f[i, s] = fn(*fin, rho, *v)
        nv = [rho * v[d] - self.mom[d](*f[:, s]) for d in range(D)] + [np.zeros_like(rho)] * (3 - D)
        for (kind, k), fn in self.zh_fix[o]:
            val = fn(*f[:, s], *nv) * np.ones_like(rho)
            if kind == 'nv':
                nv[k] = val
            else:
                f[k, s] = val

    def _bgk(self, f, rho, v, s):
        """BGK_relaxate0, relaxation.mako:99-181."""
        D, Q = self.D, self.Q
        v0 = v[:, s].copy()
        if self.accel is not None:
            v0 += 0.5 * self.accel[:, None]
        omega = 1.0 / self.tau
        for i in range(Q):
            feq = self.feq[i](rho[s], *v0)
            f[i, s] += omega * (feq - f[i, s])
        if self.accel is not None:
            pref = self.guo_pref(rho[s], self.tau)
            a = [self.accel[d] * np.ones(v0.shape[1]) for d in range(D)]
            for i in range(Q):
                f[i, s] += self.guo[i](pref, *v0, *a)
            v[:, s] += 0.5 * self.accel[:, None]

    def _mrt(self, f, v, s, m):
        """MS_relaxate, relaxation_mrt.mako:31-97."""
        D, Q = self.D, self.Q
        fm = self.M @ f[:, s]
        if self.accel is not None:
            for d in range(D):
                fm[self.m_idx[d]] += 0.5 * self.accel[d]
        mom = [fm[k] for k in self.m_idx]
        meq = [self.meq[k](fm[0], *mom) * np.ones(fm.shape[1]) for k in range(Q)]
        eqn = (m(nt.NTEquilibriumVelocity) | m(nt.NTEquilibriumDensity))[s]
        for k in range(Q):
            if self.rates[k] != 0:
                fm[k, eqn] = meq[k][eqn]
        for k in range(Q):
            if self.rates[k] != 0:
                fm[k] -= self.rates[k] * (fm[k] - meq[k])
        if self.accel is not None:
            for d in range(D):
                fm[self.m_idx[d]] += 0.5 * self.accel[d]
        f[:, s] = self.Minv @ fm
        if self.accel is not None:
            v[:, s] += 0.5 * self.accel[:, None]

    def _missing_links(self, ori, s):
        """for_orientation_dist(link_tags=True), boundary.mako:391-416: yields (node selection, missing_idx,
        opposite idx) over the nodes selected by `s`."""
        g, D = self.grid, self.D
        if self.ctx.get('use_link_tags', True):
            unused = int(self.ctx.get('unused_tag_bits', 0))
            for k in range(1, self.Q):
                if unused & (1 << (k - 1)):
                    continue
                sel = s & ((ori & (1 << (k - 1))) == 0)
                if sel.any():
                    yield sel, g.idx_opposite[k], k
        else:
            for on in range(1, 2 * D + 1):
                sel = s & (ori == on)
                if sel.any():
                    for k in self.missing[on]:
                        yield sel, k, g.idx_opposite[k]

    def _halfbb(self, f, dout, idx, ori, s, odd):
        """_global_mem_fill_missing_dists_with_opposites, boundary.mako:653-668."""
        if not s.any():
            return
        for sel, missing, name in self._missing_links(ori, s):
            self._hbb_write(f[name, sel], dout, tuple(a[sel] for a in idx), missing, odd)

    def _hbb_write(self, vals, dout, sub, missing, odd):
        g, D = self.grid, self.D
        if (not self.aa) or odd:
            dout[missing][sub] = vals
        else:
            oi = g.idx_opposite[missing]
            e = [int(c) for c in g.basis[oi]]
            tgt = tuple(sub[k] + e[D - 1 - k] for k in range(D))
            dout[oi][tgt] = vals

    def _pbc(self, dist, swap):
        """Semantic statement of ApplyPeriodicBoundaryConditions[WithSwap]
        (kernel_utils.mako:19-384): axis passes in x, y, z order; every pass moves the
        populations that crossed (or, for the AA in-place step, will be read across)
        the periodic face over the whole face, other-axis ghosts included, so that
        edge and corner crossings compose.  Non-finite values are never copied."""
        g, D, Q = self.grid, self.D, self.Q
        for axis in range(D):
            if not self.periodic[axis]:
                continue
            ax = D - 1 - axis          # numpy axis
            n = self.lat[ax]

            def sl(c):
                s = [slice(None)] * D
                s[ax] = c
                s[D - 1] = slice(0, self.lat[-1]) if ax != D - 1 else c
                return tuple(s)
            for i in range(1, Q):
                e = int(g.basis[i][axis])
                if e == 0:
                    continue
                if not swap:
                    # populations that left through the face sit in the ghost layer
                    src, dst = (0, n - 2) if e == -1 else (n - 1, 1)
                    slot = i
                else:
                    # unpropagated data in opposite slots: real layer -> far ghost layer
                    src, dst = (1, n - 1) if e == -1 else (n - 2, 0)
                    slot = g.idx_opposite[i]
                a = dist[slot][sl(src)]
                b = dist[slot][sl(dst)]
                ok = np.isfinite(a)
                if not swap:
                    # A diagonal population cannot have arrived at the first/last real
                    # row of a NON-periodic other axis from inside the domain (make_cond,
                    # kernel_utils.mako:120-171): those positions are skipped.
                    for oaxis in range(D):
                        eo = int(g.basis[i][oaxis])
                        if oaxis == axis or eo == 0 or self.periodic[oaxis]:
                            continue   # periodic other axes: the later axis pass completes the wrap
                        oax = D - 1 - oaxis
                        coord = np.arange(self.lat[oax])
                        good = coord >= 2 if eo == 1 else coord <= self.lat[oax] - 3
                        shp = [1] * (D - 1)
                        # position of oax inside the (D-1)-dim face array
                        pos = oax if oax < ax else oax - 1
                        shp[pos] = good.size
                        ok = ok & good.reshape(shp)
                b[ok] = a[ok]


class RefShanChenStepper(object):
    """Binary Shan-Chen mixture (lb_binary.py:375-517) with the reference's expressions:
    equilibria = LBBinaryFluidShanChen.equilibria (bgk_equilibrium for rho and for phi),
    pseudopotential = sym.SHAN_CHEN_POTENTIALS, forcing = sym_force.guo_external_force{,_pref}
    for grid 0 and 1.  Control flow: binary_shan_chen.mako:19-141, shan_chen.mako:9-84,
    NNSubdomainRunner.step (subdomain_runner.py:2102-2197).  fp64, all-wet or FullBB geometries."""

    def __init__(self, runner, ctx):
        self.base = RefStepper.__new__(RefStepper)
        b = self.base
        # reuse geometry decoding of RefStepper
        RefStepper.__init__(b, runner, dict(ctx, model='bgk', forces=None))
        g, S, cfg = b.grid, sym.S, b.cfg
        self.g, self.D, self.Q = g, b.D, b.Q
        self.aa = b.aa
        self.ctx = ctx
        self.tau = [float(ctx['tau']), float(ctx['tau_phi'])]
        consts = ctx['constants']
        self.G = [[consts['G11'], consts['G12']], [consts['G21'], consts['G22']]]
        x = sympy.Symbol('x')
        self.ppot = _lam([x], sym.SHAN_CHEN_POTENTIALS[ctx['sc_potential']]('x'))
        vs = b.vs
        eq0 = sym_equilibrium.bgk_equilibrium(g, cfg).expression
        eq1 = sym_equilibrium.bgk_equilibrium(g, cfg, rho=S.phi, rho0=S.phi).expression
        self.feq = [[_lam([S.rho] + vs, e) for e in eq0], [_lam([S.phi] + vs, e) for e in eq1]]
        pref = sympy.Symbol('pref')
        self.guo, self.guo_pref = [], []
        for gn, dens, tsym in ((0, S.rho, S.tau0), (1, S.phi, S.tau1)):
            ea = [getattr(S, 'g%dea%s' % (gn, c)) for c in 'xyz'][:self.D]
            self.guo.append([_lam([pref] + vs + ea, e) for e in sym_force.guo_external_force(g, grid_num=gn)])
            self.guo_pref.append(_lam([dens, tsym], sym_force.guo_external_force_pref(g, cfg, grid_num=gn)))
        self.rho_std, self.mom = b.rho_std, b.mom
        shape = b.shape
        self.d = [[np.zeros((self.Q,) + shape), None if self.aa else np.zeros((self.Q,) + shape)] for _ in range(2)]
        self.iteration = 0

    def set_initial_conditions(self, rho, phi, v):
        b = self.base
        with np.errstate(all='ignore'):
            okmask = b.wet | b.is_cls(nt.NTFullBBWall) if nt.NTFullBBWall in self.ctx['node_types'] else b.wet
            for gidx, dens in ((0, rho), (1, phi)):
                for i in range(self.Q):
                    val = self.feq[gidx][i](dens, *v) * np.ones(b.shape)
                    val[~okmask] = np.inf
                    val[..., b.lat[-1]:] = 0.0
                    for arr in self.d[gidx]:
                        if arr is not None:
                            arr[i] = val
        self.rho, self.phi = np.array(rho, float).copy(), np.array(phi, float).copy()
        self.v = [np.array(c, float).copy() for c in v]

    def _io(self, gidx):
        a, bb = self.d[gidx]
        if bb is None:
            return a, a
        return (a, bb) if (self.iteration & 1) == 0 else (bb, a)

    def current(self, gidx):
        a, bb = self.d[gidx]
        return a if (bb is None or (self.iteration & 1) == 0) else bb

    def _load(self, din, idx, odd):
        g, D, Q = self.g, self.D, self.Q
        f = np.zeros((Q, idx[0].size))
        for i in range(Q):
            if not odd:
                f[i] = din[i][idx]
            else:
                e = [int(c) for c in g.basis[i]]
                f[i] = din[g.idx_opposite[i]][tuple(idx[k] - e[D - 1 - k] for k in range(D))]
        return f

    def step(self):
        b, g, D, Q = self.base, self.g, self.D, self.Q
        it = self.iteration
        odd = self.aa and (it & 1)
        # ShanChenPrepareMacroFields: wet nodes only
        idx = np.nonzero(b.active & b.wet)
        ti = [1.0 / t for t in self.tau]
        f0 = self._load(self._io(0)[0], idx, odd)
        f1 = self._load(self._io(1)[0], idx, odd)
        rho0, rho1 = self.rho_std(*f0), self.rho_std(*f1)
        total = ti[0] * rho0 + ti[1] * rho1
        self.rho[idx], self.phi[idx] = rho0, rho1
        for d in range(D):
            self.v[d][idx] = (ti[0] * self.mom[d](*f0) + ti[1] * self.mom[d](*f1)) / total
        # ApplyMacroPeriodicBoundaryConditions (kernel_utils.mako:410-474), x, y, z passes
        for field in (self.rho, self.phi):
            for axis in range(D):
                if not b.periodic[axis]:
                    continue
                ax = D - 1 - axis
                n = b.lat[ax]
                sl = lambda c: tuple(c if k == ax else (slice(0, b.lat[-1]) if k == D - 1 else slice(None)) for k in range(D))
                src = field[sl(n - 2)]; dst = field[sl(0)]; ok = np.isfinite(src); dst[ok] = src[ok]
                src = field[sl(1)]; dst = field[sl(n - 1)]; ok = np.isfinite(src); dst[ok] = src[ok]
        # ShanChenCollideAndPropagate{0,1}
        idx = np.nonzero(b.active)
        N = idx[0].size
        wet = b.wet[idx]
        fields = [self.rho, self.phi]
        wall = b.is_cls(nt.NTFullBBWall)[idx] if nt.NTFullBBWall in self.ctx['node_types'] else np.zeros(N, bool)
        for gidx in range(2):
            din, dout = self._io(gidx)
            gm = fields[gidx][idx]
            acc = np.zeros((D, N))
            for k in range(2):
                Gc = self.G[gidx][k]
                if Gc == 0.0:
                    continue
                force = np.zeros((D, N))
                for i in range(1, Q):
                    e = [int(c) for c in g.basis[i]]
                    nb = tuple(idx[kk] + e[D - 1 - kk] for kk in range(D))
                    psi = self.ppot(fields[k][nb])
                    for d in range(D):
                        if e[d]:
                            force[d] += psi * (e[d] * float(g.weights[i]))
                acc += force * (-self.ppot(gm) * Gc)
            with np.errstate(all='ignore'):
                acc = np.where(wet, acc / gm, 0.0)
            # body_force(), relaxation_common.mako:24-33: ea<grid>[j] += body_force_accel(grid, j, forces)
            forces = self.ctx.get('forces')
            if forces is not None and getattr(forces, 'numeric', None) and gidx in forces.numeric:
                for d in range(D):
                    acc[d] += np.where(wet, float(sym_force.body_force_accel(gidx, d, forces, accel=True)), 0.0)
            f = self._load(din, idx, odd)
            v = np.array([self.v[d][idx] for d in range(D)])
            if wall.any():
                for i in sym.bb_swap_pairs(g):
                    j = g.idx_opposite[i]
                    t = f[i, wall].copy(); f[i, wall] = f[j, wall]; f[j, wall] = t
            s = wet
            v0 = v[:, s] + 0.5 * acc[:, s]
            omega = 1.0 / self.tau[gidx]
            for i in range(Q):
                f[i, s] += omega * (self.feq[gidx][i](gm[s], *v0) - f[i, s])
            pref = self.guo_pref[gidx](gm[s], self.tau[gidx])
            for i in range(Q):
                f[i, s] += self.guo[gidx][i](pref, *v0, *acc[:, s])
            if self.aa and not odd:
                for i in range(Q):
                    dout[g.idx_opposite[i]][idx] = f[i]
            else:
                for i in range(Q):
                    e = [int(c) for c in g.basis[i]]
                    ok = np.ones(N, bool)
                    for k in range(D):
                        c = e[D - 1 - k]
                        if c > 0:
                            ok &= idx[k] < b.lat[k] - 1
                        elif c < 0:
                            ok &= idx[k] > 0
                    dout[i][tuple(idx[k][ok] + e[D - 1 - k] for k in range(D))] = f[i, ok]
            b._pbc(dout, swap=(self.aa and not odd))
        self.iteration += 1
