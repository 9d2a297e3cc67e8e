"""Drives several subdomains with the CPU oracle the way the B200 runner drives them on GPUs
(sailfish_b200/subdomain_runner.py step()): collide, local PBC, then the planes of
tests/halo_ref.face_transfer move into the neighbours' lattices."""
import numpy as np

import halo_ref as halo
from sailfish_b200 import params as P
from sailfish_b200.controller import LBSimulationController


def build_case(cls, geo, args):
    ctrl = LBSimulationController(cls, geo)
    cfg = ctrl._parser.parse(args)
    ctrl.config = cfg
    cls.modify_config(cfg)
    specs = ctrl._make_specs()
    parts = []
    for spec in specs:
        sim = cls(cfg)
        dim = spec.dim
        gsize = (cfg.lat_ny, cfg.lat_nx) if dim == 2 else (cfg.lat_nz, cfg.lat_ny, cfg.lat_nx)
        sd = cls.subdomain(gsize, spec, sim.grid, cfg)
        sd.reset()
        base = []
        for _ in range(1 + dim):
            f = np.full(sd.arr_shape, np.inf, dtype=np.float32)
            f[spec._nonghost_slice] = 0
            base.append(f)
        sim.rho = base[0][spec._nonghost_slice]
        sim.v = [b[spec._nonghost_slice] for b in base[1:]]
        for i, c in enumerate('xyz'[:dim]):
            setattr(sim, 'v' + c, sim.v[i])
        sd.init_fields(sim)
        lat = list(reversed(sd.lat_shape)) + ([1] if dim == 2 else [])
        arr = list(reversed(sd.arr_shape)) + ([1] if dim == 2 else [])
        per = list(spec._periodicity) + [False] * (3 - dim)
        p = P.make_params(sim.grid.name, cfg.model, cfg.access_pattern, lat, arr, per, sd.bits,
                          sd.type_ids_by_name(), sd.node_params, cfg.visc, sim.body_acceleration(),
                          use_link_tags=cfg.use_link_tags,
                          relaxation_enabled=getattr(cfg, 'relaxation_enabled', True),
                          propagation_enabled=getattr(cfg, 'propagation_enabled', True))
        parts.append(dict(spec=spec, sd=sd, sim=sim, params=p, rho=base[0], v=base[1:], grid=sim.grid))
    return cfg, parts


class OracleSubdomain(object):
    def __init__(self, oracle, part, dtype):
        self.part = part
        self.o = oracle.Simulation(part['params'], part['sd'].encoded_map(), part['rho'], part['v'], dtype=dtype)
        self.oracle = oracle
        self.shape = part['sd'].arr_shape
        self.Q = part['grid'].Q

    def dist_view(self, which):
        return which.reshape((self.Q,) + self.shape)

    def collide_and_pbc(self, save):
        """Everything SubdomainRunner.step does before the planes leave."""
        s = self.o
        it = s.iteration
        if s.dist_b is None:
            din = dout = s.dist_a
        elif (it & 1) == 0:
            din, dout = s.dist_a, s.dist_b
        else:
            din, dout = s.dist_b, s.dist_a
        self.oracle.collide_and_propagate(s.p, s.map, din, dout, s.rho, s.v, 1 if save else 0, it)
        aa = s.dist_b is None
        for axis in range(s.p.dim):
            if s.p.periodic[axis]:
                self.oracle.apply_pbc(s.p, dout, axis, with_swap=(aa and (it & 1) == 0))
        self.dout = dout
        self.swap = halo.with_swap_for(aa, it)

    def outgoing(self, face, peer_lat_n):
        slots, src, dst = halo.face_transfer(self.part['grid'], face, self.swap, self.shape[0], peer_lat_n)
        return slots, dst, halo.extract_planes(self.dist_view(self.dout), slots, src)

    def incoming(self, slots, dst, planes):
        halo.insert_planes(self.dist_view(self.dout), slots, dst, planes)

    def finish(self):
        self.o.iteration += 1


def run_local(oracle, parts, steps, dtype=np.float64):
    """All subdomains in one process."""
    subs = {p['spec'].id: OracleSubdomain(oracle, p, dtype) for p in parts}
    for k in range(steps):
        for s in subs.values():
            s.collide_and_pbc(save=(k == steps - 1))
        msgs = []
        for sid, s in subs.items():
            for face, other in s.part['spec']._face_conn.items():
                msgs.append((other,) + s.outgoing(face, subs[other].shape[0]))
        for other, slots, dst, planes in msgs:
            subs[other].incoming(slots, dst, planes)
        for s in subs.values():
            s.finish()
    return subs


def merge_field(parts, subs, getter, dim):
    """Real-node field of the global domain assembled from the subdomains."""
    cfg_shape = None
    pieces = []
    for p in parts:
        spec = p['spec']
        arr = getter(subs[spec.id]).reshape(p['sd'].arr_shape)[spec._nonghost_slice]
        pieces.append((spec.location[dim - 1], arr))
    pieces.sort(key=lambda t: t[0])
    return np.concatenate([a for _, a in pieces], axis=0)
