"""Geometry + whole-step golden fixtures (see tools/gen_golden.py).  Build container only."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
import numpy as np  # noqa: E402
from ref_stepper import RefStepper  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _periodic_box_classes():
    """All-fluid periodic box with a smooth initial field (the bandwidth microbenchmark
    of SURVEY.md 8d), written against the reference API."""
    from sailfish.subdomain import Subdomain2D, Subdomain3D
    from sailfish.lb_single import LBFluidSim

    class Box3D(Subdomain3D):
        def boundary_conditions(self, hx, hy, hz):
            pass

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 1e-3 * np.sin(2 * np.pi * hx / self.gx) * np.cos(2 * np.pi * hz / self.gz)
            sim.vx[:] = 0.05 * np.sin(2 * np.pi * hy / self.gy)
            sim.vy[:] = 0.03 * np.cos(2 * np.pi * hz / self.gz)
            sim.vz[:] = 0.02 * np.sin(2 * np.pi * (hx + hy) / self.gx)

    class BoxSim3D(LBFluidSim):
        subdomain = Box3D

        @classmethod
        def update_defaults(cls, defaults):
            defaults.update({'lat_nx': 10, 'lat_ny': 9, 'lat_nz': 8, 'grid': 'D3Q19',
                             'periodic_x': True, 'periodic_y': True, 'periodic_z': True})

    class Box2D(Subdomain2D):
        def boundary_conditions(self, hx, hy):
            pass

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + 1e-3 * np.sin(2 * np.pi * hx / self.gx)
            sim.vx[:] = 0.05 * np.sin(2 * np.pi * hy / self.gy)
            sim.vy[:] = 0.03 * np.cos(2 * np.pi * hx / self.gx)

    class BoxSim2D(LBFluidSim):
        subdomain = Box2D

        @classmethod
        def update_defaults(cls, defaults):
            defaults.update({'lat_nx': 12, 'lat_ny': 10, 'grid': 'D2Q9',
                             'periodic_x': True, 'periodic_y': True})
    return BoxSim2D, BoxSim3D


def _cases():
    ldc2 = ref_harness.load_example('examples/ldc_2d.py').LDCSim
    ldc3 = ref_harness.load_example('examples/ldc_3d.py').LDCSim
    poi3 = ref_harness.load_example('examples/poiseuille_3d.py')
    poi2 = ref_harness.load_example('examples/poiseuille.py')
    from sailfish.geo import EqualSubdomainsGeometry3D
    box2, box3 = _periodic_box_classes()
    c = []
    for ap in ('AB', 'AA'):
        c.append(('ldc2d_bgk_' + ap, ldc2, None, ['--lat_nx=16', '--lat_ny=14', '--visc=0.05', '--access_pattern=' + ap], 20))
        c.append(('ldc2d_mrt_' + ap, ldc2, None, ['--lat_nx=16', '--lat_ny=14', '--visc=0.05', '--model=mrt', '--access_pattern=' + ap], 20))
        c.append(('ldc3d_bgk_' + ap, ldc3, None, ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04', '--access_pattern=' + ap], 20))
        c.append(('ldc3d_mrt_' + ap, ldc3, None, ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04', '--model=mrt', '--access_pattern=' + ap], 20))
        c.append(('ldc3d_q27_' + ap, ldc3, None, ['--lat_nx=10', '--lat_ny=10', '--lat_nz=9', '--visc=0.04', '--grid=D3Q27', '--access_pattern=' + ap], 16))
        c.append(('poi3d_press_mrt_' + ap, poi3.PoiseuilleSim, EqualSubdomainsGeometry3D,
                  ['--grid=D3Q19', '--lat_nx=12', '--lat_ny=12', '--lat_nz=16', '--flow_direction=z', '--drive=pressure',
                   '--stationary', '--model=mrt', '--visc=0.1', '--access_pattern=' + ap], 20))
        c.append(('poi3d_force_bgk_' + ap, poi3.PoiseuilleSim, EqualSubdomainsGeometry3D,
                  ['--grid=D3Q19', '--lat_nx=14', '--lat_ny=12', '--lat_nz=12', '--flow_direction=x', '--drive=force',
                   '--visc=0.1', '--access_pattern=' + ap], 20))
        c.append(('poi3d_force_mrt_' + ap, poi3.PoiseuilleSim, EqualSubdomainsGeometry3D,
                  ['--grid=D3Q19', '--lat_nx=12', '--lat_ny=12', '--lat_nz=14', '--flow_direction=z', '--drive=force',
                   '--model=mrt', '--visc=0.1', '--access_pattern=' + ap], 20))
        c.append(('box3d_bgk_' + ap, box3, None, ['--visc=0.02', '--access_pattern=' + ap], 12))
        c.append(('box3d_q27_' + ap, box3, None, ['--visc=0.02', '--grid=D3Q27', '--access_pattern=' + ap], 12))
        c.append(('box2d_bgk_' + ap, box2, None, ['--visc=0.02', '--access_pattern=' + ap], 12))
    return c, poi2


def pack_ctx(ctx):
    remap = ctx['type_id_remap']
    names = sorted(c.__name__ for c in ctx['node_types'])
    ids = [int(remap[c.id]) for c in sorted(ctx['node_types'], key=lambda c: c.__name__)]
    accel = np.zeros(3)
    forces = ctx.get('forces')
    if forces is not None and getattr(forces, 'numeric', None) and 0 in forces.numeric:
        a = forces.numeric[0][True]
        accel[:len(a)] = a
    dim = ctx['grid'].dim
    return dict(
        grid=ctx['grid'].__name__, model=ctx['model'], access_pattern=ctx['access_pattern'],
        lat=np.array([ctx['lat_nx'], ctx['lat_ny'], ctx['lat_nz'] if dim == 3 else 1]),
        arr=np.array([ctx['arr_nx'], ctx['arr_ny'], ctx['arr_nz'] if dim == 3 else 1]),
        periodic=np.array([int(bool(x)) for x in ctx['block_periodicity']]),
        bits=np.array([ctx['nt_misc_shift'], ctx['nt_param_shift'], ctx['nt_scratch_shift']]),
        type_names=np.array(names), type_ids=np.array(ids),
        node_params=np.array([float(x) for x in ctx['node_params']] or [0.0]),
        visc=np.float64(ctx['visc']), tau=np.float64(ctx['tau']), accel=accel,
        use_link_tags=np.int32(bool(ctx.get('use_link_tags', True))))


def main(what, cases=None):
    if cases is None:
        cases, _ = _cases()
    for name, cls, geo, args, steps in cases:
        cfg, runners = ref_harness.make_runners(cls, geo, args=args)
        r = runners[0]
        ctx = ref_harness.context_of(r)
        out = pack_ctx(ctx)
        out['map'] = r._subdomain.encoded_map().astype(np.uint32)
        sim = r._sim
        rho0 = r.field_base(sim.rho).astype(np.float64)
        v0 = [r.field_base(c).astype(np.float64) for c in sim.v]
        out['rho0'] = rho0
        out['v0'] = np.array(v0)
        out['fluid_map'] = r._subdomain.fluid_map()
        if 'steps' in what:
            st = RefStepper(r, ctx)
            st.set_initial_conditions(rho0, v0)
            for k in range(steps):
                st.step(save_macro=(k == steps - 1))
            out['steps'] = np.int32(steps)
            out['dist'] = st.current_dist().copy()
            out['rho'] = st.rho
            out['v'] = np.array(st.v)
            # mass over real fluid nodes, for the record
            fin = np.isfinite(st.rho)
            if not fin.any():
                print('%-22s steps=%d  NO FINITE DENSITY -- case dropped' % (name, steps))
                continue
            print('%-22s steps=%d  rho[min,max]=%.6f %.6f  max|v|=%.4e' % (
                name, steps, np.nanmin(st.rho[fin]), np.nanmax(st.rho[fin]),
                max(np.nanmax(np.abs(c[np.isfinite(c)])) for c in st.v)))
        np.savez_compressed(os.path.join(GOLD, 'steps_%s.npz' % name), **out)





def gen_split_geometry():
    """Encoded maps of every subdomain of z-split runs (geometry only): pins the mirror's
    ghost / orientation handling at connected faces (subdomain.py:978-1005)."""
    from sailfish.geo import EqualSubdomainsGeometry3D, EqualSubdomainsGeometry2D
    ldc3 = ref_harness.load_example('examples/ldc_3d.py').LDCSim
    ldc2 = ref_harness.load_example('examples/ldc_2d.py').LDCSim
    poi3 = ref_harness.load_example('examples/poiseuille_3d.py').PoiseuilleSim
    cases = [
        ('split_ldc3d_z3', ldc3, EqualSubdomainsGeometry3D,
         ['--lat_nx=10', '--lat_ny=9', '--lat_nz=14', '--subdomains=3', '--conn_axis=z', '--access_pattern=AA']),
        ('split_ldc2d_y2', ldc2, EqualSubdomainsGeometry2D,
         ['--lat_nx=12', '--lat_ny=11', '--subdomains=2', '--conn_axis=y', '--access_pattern=AA']),
        ('split_poi3d_z2', poi3, EqualSubdomainsGeometry3D,
         ['--grid=D3Q19', '--lat_nx=10', '--lat_ny=10', '--lat_nz=12', '--flow_direction=z', '--drive=force',
          '--subdomains=2', '--conn_axis=z', '--access_pattern=AA']),
    ]
    for name, cls, geo, args in cases:
        cfg, runners = ref_harness.make_runners(cls, geo, args=args)
        out = {'n': np.int32(len(runners))}
        for i, r in enumerate(runners):
            ctx = ref_harness.context_of(r)
            pk = pack_ctx(ctx)
            for k, v in pk.items():
                out['%s_%d' % (k, i)] = v
            out['map_%d' % i] = r._subdomain.encoded_map().astype(np.uint32)
            out['location_%d' % i] = np.array(r._spec.location)
            out['size_%d' % i] = np.array(r._spec.size)
            out['fluid_map_%d' % i] = r._subdomain.fluid_map()
        np.savez_compressed(os.path.join(GOLD, 'geometry_%s.npz' % name), **out)
        print('geometry_%s.npz' % name, [tuple(out['size_%d' % i]) for i in range(len(runners))])



def gen_shan_chen(only=None):
    """Binary Shan-Chen goldens (BASELINE config 5, examples/binary_fluid/sc_separation_3d.py)."""
    from ref_stepper import RefShanChenStepper
    ex3 = ref_harness.load_example('examples/binary_fluid/sc_separation_3d.py')
    from sailfish.subdomain import Subdomain3D, Subdomain2D
    from sailfish.lb_binary import LBBinaryFluidShanChen
    from sailfish.lb_base import LBForcedSim
    from sailfish.node_type import NTFullBBWall
    from sailfish.geo import LBGeometry3D, LBGeometry2D

    class WalledDomain(Subdomain3D):
        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 0.01 * np.sin(2 * np.pi * hx / self.gx) * np.cos(2 * np.pi * hy / self.gy)
            sim.phi[:] = 1.0 - 0.01 * np.sin(2 * np.pi * hx / self.gx) * np.cos(2 * np.pi * hy / self.gy)

        def boundary_conditions(self, hx, hy, hz):
            self.set_node((hz == 0) | (hz == self.gz - 1), NTFullBBWall)

    class WalledSim(LBBinaryFluidShanChen):
        subdomain = WalledDomain

        @classmethod
        def update_defaults(cls, defaults):
            defaults.update({'lat_nx': 10, 'lat_ny': 9, 'lat_nz': 8, 'grid': 'D3Q19', 'G12': 0.9, 'G11': -0.3,
                             'visc': 0.12, 'tau_phi': 0.9, 'periodic_x': True, 'periodic_y': True})

    class Sep2D(Subdomain2D):
        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + np.random.rand(*sim.rho.shape) / 1000.0
            sim.phi[:] = 1.0 + np.random.rand(*sim.phi.shape) / 1000.0

        def boundary_conditions(self, hx, hy):
            pass

    class Sep2DSim(LBBinaryFluidShanChen):
        subdomain = Sep2D

        @classmethod
        def update_defaults(cls, defaults):
            defaults.update({'lat_nx': 14, 'lat_ny': 12, 'grid': 'D2Q9', 'G12': 1.1, 'visc': 1.0 / 6.0,
                             'periodic_x': True, 'periodic_y': True, 'sc_potential': 'classic'})

    class GravityDomain2D(Subdomain2D):
        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + 0.02 * np.sin(2 * np.pi * hx / self.gx) * np.cos(np.pi * hy / self.gy)
            sim.phi[:] = 1.0 - 0.02 * np.sin(2 * np.pi * hx / self.gx) * np.cos(np.pi * hy / self.gy)

        def boundary_conditions(self, hx, hy):
            self.set_node((hy == 0) | (hy == self.gy - 1), NTFullBBWall)

    class GravitySim2D(LBBinaryFluidShanChen, LBForcedSim):
        """Both components pushed along x between two walls (examples/binary_fluid/sc_poiseuille_2d.py)."""
        subdomain = GravityDomain2D

        @classmethod
        def update_defaults(cls, defaults):
            defaults.update({'lat_nx': 16, 'lat_ny': 12, 'grid': 'D2Q9', 'G12': 0.9, 'visc': 0.12, 'tau_phi': 0.8,
                             'periodic_x': True})

        def __init__(self, config):
            super(GravitySim2D, self).__init__(config)
            self.add_body_force((1e-4, 0.0), grid=0)
            self.add_body_force((2e-4, -5e-5), grid=1)

    class GravitySim3D(WalledSim, LBForcedSim):
        """Only the second component feels the force (examples/binary_fluid/sc_rayleigh_taylor_2d.py)."""
        def __init__(self, config):
            super(GravitySim3D, self).__init__(config)
            self.add_body_force((1e-4, 0.0, -2e-4), grid=1)

    cases = []
    for ap in ('AB', 'AA'):
        if only and 'gravity' in only:
            cases.append(('sc2d_gravity_' + ap, GravitySim2D, LBGeometry2D, ['--seed=1234', '--access_pattern=' + ap], 12))
            cases.append(('sc3d_gravity_' + ap, GravitySim3D, LBGeometry3D, ['--seed=1234', '--access_pattern=' + ap], 12))
            continue
        cases.append(('sc3d_' + ap, ex3.SeparationSCSim, LBGeometry3D,
                      ['--lat_nx=12', '--lat_ny=10', '--lat_nz=9', '--seed=1234', '--access_pattern=' + ap], 12))
        cases.append(('sc3d_walls_' + ap, WalledSim, LBGeometry3D, ['--seed=1234', '--access_pattern=' + ap], 12))
        cases.append(('sc2d_classic_' + ap, Sep2DSim, LBGeometry2D, ['--seed=1234', '--access_pattern=' + ap], 12))
    for name, cls, geo, args, steps in cases:
        cfg, runners = ref_harness.make_runners(cls, geo, args=args)
        r = runners[0]
        ctx = ref_harness.context_of(r)
        out = pack_ctx(ctx)
        out['simtype'] = 'shan-chen'
        out['tau_phi'] = np.float64(ctx['tau_phi'])
        c = ctx['constants']
        out['sc_g'] = np.array([c['G11'], c['G12'], c['G21'], c['G22']], dtype=np.float64)
        out['sc_potential'] = ctx['sc_potential']
        a1 = np.zeros(3)
        fr = ctx.get('forces')
        if fr is not None and getattr(fr, 'numeric', None) and 1 in fr.numeric:
            a1[:len(fr.numeric[1][True])] = fr.numeric[1][True]
        out['accel1'] = a1
        out['map'] = r._subdomain.encoded_map().astype(np.uint32)
        sim = r._sim
        rho0 = r.field_base(sim.rho).astype(np.float64)
        phi0 = r.field_base(sim.phi).astype(np.float64)
        v0 = [r.field_base(cc).astype(np.float64) for cc in sim.v]
        out['rho0'], out['phi0'], out['v0'] = rho0, phi0, np.array(v0)
        out['fluid_map'] = r._subdomain.fluid_map()
        st = RefShanChenStepper(r, ctx)
        st.set_initial_conditions(rho0, phi0, v0)
        for k in range(steps):
            st.step()
        out['steps'] = np.int32(steps)
        out['dist1'], out['dist2'] = st.current(0).copy(), st.current(1).copy()
        out['rho'], out['phi'], out['v'] = st.rho, st.phi, np.array(st.v)
        fm = np.zeros(st.rho.shape, bool)
        fm[r._spec._nonghost_slice] = out['fluid_map']
        print('%-20s steps=%d rho[%.6f %.6f] phi[%.6f %.6f] max|v|=%.3e mass=%.10f' % (
            name, steps, st.rho[fm].min(), st.rho[fm].max(), st.phi[fm].min(), st.phi[fm].max(),
            max(np.abs(cc[fm]).max() for cc in st.v), st.rho[fm].sum() + st.phi[fm].sum()))
        np.savez_compressed(os.path.join(GOLD, 'steps_%s.npz' % name), **out)


def gen_geometry_zoo():
    """tests/geometry_zoo.py run with the reference's classes: encoded maps, parameters, bit layout, remap."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import geometry_zoo
    from sailfish import node_type as rnt
    from sailfish.subdomain import Subdomain2D, Subdomain3D
    from sailfish.lb_single import LBFluidSim
    from sailfish.lb_base import LBForcedSim
    from sailfish.geo import LBGeometry2D, LBGeometry3D
    ns = dict(Subdomain2D=Subdomain2D, Subdomain3D=Subdomain3D, LBFluidSim=LBFluidSim, LBForcedSim=LBForcedSim)
    for n in dir(rnt):
        if n.startswith('NT'):
            ns[n] = getattr(rnt, n)
    out = {}
    for name, (cls, args) in sorted(geometry_zoo.make_cases(ns).items()):
        for ap in ('AB', 'AA'):
            d = {}
            cls.update_defaults(d)
            geo = LBGeometry2D if d['grid'] == 'D2Q9' else LBGeometry3D
            cfg, runners = ref_harness.make_runners(cls, geo, args=args + ['--access_pattern=' + ap])
            r = runners[0]
            ctx = ref_harness.context_of(r)
            pk = pack_ctx(ctx)
            key = '%s_%s' % (name, ap)
            out[key + '/map'] = r._subdomain.encoded_map().astype(np.uint32)
            out[key + '/bits'] = pk['bits']
            out[key + '/node_params'] = pk['node_params']
            out[key + '/type_names'] = pk['type_names']
            out[key + '/type_ids'] = pk['type_ids']
            out[key + '/fluid_map'] = r._subdomain.fluid_map()
            out[key + '/unused_tag_bits'] = np.int64(ctx.get('unused_tag_bits', 0))
            out[key + '/rho0'] = r.field_base(r._sim.rho).astype(np.float32)
            print('%-24s map %s types %s' % (key, out[key + '/map'].shape, dict(zip(pk['type_names'], pk['type_ids']))))
    np.savez_compressed(os.path.join(GOLD, 'geometry_zoo.npz'), **out)


def gen_zoo_steps():
    """Step goldens on the geometry zoo: pins half-way bounce-back, equilibrium-velocity, regularised-density and
    do-nothing nodes (the examples only use full-way walls, regularised velocity and equilibrium density)."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'tests'))
    import geometry_zoo
    from sailfish import node_type as rnt
    from sailfish.subdomain import Subdomain2D, Subdomain3D
    from sailfish.lb_single import LBFluidSim
    from sailfish.lb_base import LBForcedSim
    from sailfish.geo import LBGeometry2D, LBGeometry3D
    ns = dict(Subdomain2D=Subdomain2D, Subdomain3D=Subdomain3D, LBFluidSim=LBFluidSim, LBForcedSim=LBForcedSim)
    for n in dir(rnt):
        if n.startswith('NT'):
            ns[n] = getattr(rnt, n)
    cases = []
    for name, (cls, args) in sorted(geometry_zoo.make_cases(ns).items()):
        d = {}
        cls.update_defaults(d)
        geo = LBGeometry2D if d['grid'] == 'D2Q9' else LBGeometry3D
        for ap in ('AB', 'AA'):
            # half-way walls on the outermost layer (no unused layer behind them) are the configuration the
            # reference warns about; without link tags even its AA == AB invariant does not hold there, and with
            # them only the AB run is used as a pin
            # without link tags a wall node's bounced populations and its wall neighbours' propagated ones target the
            # same slots (a write race in the reference's kernel): geometry goldens only
            if name in ('duct3d_notags', 'tms_channel2d_notags'):
                continue
            if ap == 'AA' and (name in geometry_zoo.AB_ONLY or name in geometry_zoo.AB_PINNED_ONLY):
                continue
            cases.append(('zoo_%s_%s' % (name, ap), cls, geo, args + ['--access_pattern=' + ap, '--visc=0.06'], 10))
    main(['steps'], cases)


def gen_slip_pairs():
    """sym.slip_bb_swap_pairs (sym.py:481-497) per lattice and orientation: pins the generated mirror tables of
    sailfish_b200/csrc/lattice_tables.cuh (tests/test_lattice_tables.py)."""
    from sailfish import sym
    out = {}
    for name in ('D2Q9', 'D3Q19', 'D3Q27'):
        g = getattr(sym, name)
        for o in range(1, 2 * g.dim + 1):
            out['%s_%d' % (name, o)] = np.array(sorted(sym.slip_bb_swap_pairs(g, o)), dtype=np.int64)
            out['%s_%d_normal' % (name, o)] = np.array([int(c) for c in g.dir_to_vec(o)], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, 'slip_pairs.npz'), **out)


if __name__ == '__main__':
    what = sys.argv[1:] or ['geometry', 'steps', 'split', 'sc']
    if 'geometry' in what or 'steps' in what:
        main(what)
    if 'split' in what:
        gen_split_geometry()
    if 'sc' in what:
        gen_shan_chen()
    if 'zoo' in what:
        gen_geometry_zoo()
        gen_zoo_steps()
    if 'zoo_steps' in what:
        gen_zoo_steps()
    if 'slip_pairs' in what:
        gen_slip_pairs()
    if 'sc_gravity' in what:
        gen_shan_chen(only='gravity')
