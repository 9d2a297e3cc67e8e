"""``SubdomainRunner`` for the B200 backend.

Mirror of sailfish/subdomain_runner.py for the single-fluid hot path: same constructor
role, same helper names (``make_scalar_field``, ``make_vector_field``, ``gpu_field``,
``gpu_dist``, ``gpu_geo_map``, ``step``, ``main``, ``_debug_get_dist/_debug_set_dist``)
and the same per-step order (boundary slabs -> halo -> bulk -> PBC), but

  * no run-time code generation: the context becomes an ``SlfParams`` plan,
  * the collect / D2H / ZeroMQ / H2D / distribute chain (subdomain_runner.py:1064-1139) is
    one peer-to-peer plane copy into the neighbour's lattice (``slf_halo_push``), ordered by
    device-side semaphores -- the host never touches halo data.
"""
import ctypes
import os
import time

import numpy as np

from sailfish_b200 import backend_b200 as bk
from sailfish_b200 import params as P
from sailfish_b200.subdomain import SubdomainSpec

_u64 = ctypes.c_uint64


class HaloLink(object):
    """One directed neighbour relation across the split axis.

    Same process (several subdomains driven by one process, the reference's "several
    subdomains per GPU" regtest mode): ``peer_runner`` is set and ordering uses CUDA events.
    Other rank: the neighbour's lattice and semaphore are CUDA-IPC mappings and ordering uses
    the stream-ordered device semaphores of slf_sem_signal / slf_sem_wait."""

    def __init__(self, face, peer_dist_a, peer_dist_b, peer_lat_n, peer_arr_n, peer_ready=0,
                 my_ready=0, my_consumed=0, peer_runner=None):
        self.face = face
        self.peer_dist_a = int(peer_dist_a)
        self.peer_dist_b = int(peer_dist_b) if peer_dist_b else 0
        self.peer_lat_n = int(peer_lat_n)
        self.peer_arr_n = int(peer_arr_n)
        self.peer_ready = int(peer_ready)
        self.my_ready = int(my_ready)
        self.my_consumed = int(my_consumed)
        self.peer_runner = peer_runner


class LBOutput(object):
    """Minimal mirror of sailfish/io.py:32-85 + NPYOutput (:301-348): one .npz per
    subdomain and output step, non-wet nodes masked with NaN."""

    def __init__(self, config, subdomain_id):
        self.basename = getattr(config, 'output', '')
        self.subdomain_id = subdomain_id
        self.digits = str(int(np.log10(config.max_iters)) + 1) if getattr(config, 'max_iters', 0) else '7'
        self._scalar, self._vector, self._fluid_map = {}, {}, None

    def register_field(self, field, name):
        if isinstance(field, list):
            self._vector[name] = field
        else:
            self._scalar[name] = field

    def set_fluid_map(self, fm):
        self._fluid_map = fm

    def mask_nonfluid_nodes(self):
        nf = np.logical_not(self._fluid_map)
        for f in self._scalar.values():
            f[nf] = np.nan
        for fv in self._vector.values():
            for f in fv:
                f[nf] = np.nan

    def verify(self):
        fm = self._fluid_map
        return (all(np.all(np.isfinite(f[fm])) for f in self._scalar.values()) and
                all(np.all(np.isfinite(c[fm])) for f in self._vector.values() for c in f))

    def filename(self, it):
        return ('{0}.{1}.{2:0' + self.digits + 'd}.npz').format(self.basename, self.subdomain_id, it)

    def save(self, it):
        self.mask_nonfluid_nodes()
        if not self.basename:
            return
        data = dict(self._scalar)
        for k, v in self._vector.items():
            data[k] = np.array(v)
        np.savez(self.filename(it), **data)


class SubdomainRunner(object):
    def __init__(self, simulation, spec, output=None, backend=None, quit_event=None, device=0):
        self._sim = simulation
        self._spec = spec
        spec.runner = self
        self.config = simulation.config
        self.backend = backend or bk.B200Backend(self.config, device)
        self._output = output or LBOutput(self.config, spec.id)
        self._quit_event = quit_event
        self._scalar_fields, self._vector_fields = [], []
        self._field_base = {}
        self._links = {}
        self._lib = bk.lib()
        self.float = np.float32
        if getattr(self.config, 'precision', 'single') != 'single':
            raise P.UnsupportedConfiguration('the B200 hot path is fp32 (like the reference default)')
        self._init_geometry()

    dim = property(lambda self: self._spec.dim)

    # -- geometry / fields -------------------------------------------------------------------
    def _init_geometry(self):
        cfg = self.config
        gsize = (cfg.lat_ny, cfg.lat_nx) if self.dim == 2 else (cfg.lat_nz, cfg.lat_ny, cfg.lat_nx)
        self._global_size = gsize
        self._subdomain = self._sim.subdomain(gsize, self._spec, self._sim.grid, cfg,
                                              mem_alignment=getattr(cfg, 'mem_alignment', 32))
        self._lat_size = list(self._subdomain.lat_shape)
        self._physical_size = list(self._subdomain.arr_shape)
        self._subdomain.reset()
        self._output.set_fluid_map(self._subdomain.fluid_map())

    def make_scalar_field(self, dtype=None, name=None, register=True, async_=True):
        """subdomain_runner.py:253-320: padded + ghost-inclusive backing array (inf in the
        envelope), the returned view hides ghosts."""
        dtype = dtype or self.float
        field = self.backend.alloc_async_host_buf(self._physical_size, dtype)
        if dtype == self.float:
            field[:] = np.inf
        view = field[self._spec._nonghost_slice]
        self._field_base[id(view)] = field
        view[:] = 0
        if name is not None and register:
            self._output.register_field(view, name)
        if register:
            self._scalar_fields.append(view)
        return view, None

    def make_vector_field(self, name=None):
        comps = [self.make_scalar_field(register=False)[0] for _ in range(self.dim)]
        if name is not None:
            self._output.register_field(comps, name)
        self._vector_fields.append(comps)
        return comps

    def field_base(self, view):
        return self._field_base[id(view)]

    def _init_fields(self):
        sim = self._sim
        sim.rho, _ = self.make_scalar_field(name='rho')
        sim.v = self.make_vector_field(name='v')
        for i, c in enumerate('xyz'[:self.dim]):
            setattr(sim, 'v' + c, sim.v[i])
        self._subdomain.init_fields(sim)

    def make_params(self):
        sd, cfg, g = self._subdomain, self.config, self._sim.grid
        lat = list(reversed(sd.lat_shape)) + ([1] if self.dim == 2 else [])
        arr = list(reversed(sd.arr_shape)) + ([1] if self.dim == 2 else [])
        per = list(self._spec._periodicity) + [False] * (3 - self.dim)
        return P.make_params(
            g.name, cfg.model, cfg.access_pattern, lat, arr, per, sd.bits, sd.type_ids_by_name(),
            sd.node_params, cfg.visc, self._sim.body_acceleration(),
            use_link_tags=getattr(cfg, 'use_link_tags', True),
            relaxation_enabled=getattr(cfg, 'relaxation_enabled', True),
            propagation_enabled=getattr(cfg, 'propagation_enabled', True))

    def _init_compute(self):
        from sailfish_b200.runtime import DeviceLattice
        self._params = self.make_params()
        sim = self._sim
        self._dev = DeviceLattice(self._params, self._subdomain.encoded_map(),
                                  self.field_base(sim.rho), [self.field_base(c) for c in sim.v],
                                  backend=self.backend,
                                  host_fields=(self.field_base(sim.rho), [self.field_base(c) for c in sim.v]))
        self._calc_stream = self._dev.stream
        self._data_stream = self.backend.make_stream(high_priority=True)
        # semaphores: [ready_low, ready_high, consumed_low, consumed_high]
        self._sems = self.backend.alloc_buf(size=64)
        self._push_events = [self.backend.make_event(self._data_stream), self.backend.make_event(self._data_stream)]
        self._slab_event = self.backend.make_event(self._calc_stream)

    def initialize(self):
        self._init_fields()
        self._init_compute()

    # -- reference-compatible accessors --------------------------------------------------------
    def gpu_field(self, field):
        if isinstance(field, list):
            return list(self._dev.d_v)
        return self._dev.d_rho

    def gpu_dist(self, num, copy):
        if copy == 0 or self._dev.d_dist_b is None:
            return self._dev.d_dist_a
        return self._dev.d_dist_b

    def gpu_geo_map(self):
        return self._dev.d_map

    def _debug_get_dist(self, output=True):
        """[Q, (nz+2), ny+2, arr_nx] array of the lattice most recently written."""
        return self._dev.get_dist()

    def _debug_set_dist(self, dist, output=True):
        self._dev.set_dist(dist)

    @property
    def num_fluid_nodes(self):
        return self._subdomain.num_fluid_nodes

    # -- halo ------------------------------------------------------------------------------------
    def sem_ptrs(self):
        base = int(self._sems)
        return {'ready': {0: base, 1: base + 4}, 'consumed': {0: base + 8, 1: base + 12}}

    def add_link(self, link):
        self._links[link.face & 1] = link

    def _outer_n(self):
        return self._lat_size[0]

    # -- stepping ---------------------------------------------------------------------------------
    def step(self, sync_req):
        """One time step: boundary slabs -> push planes to the neighbours (data stream) while the
        interior is computed (calc stream) -> local PBC.  Cf. subdomain_runner.py:960-1058."""
        dev = self._dev
        it = dev.iteration
        swap = int(dev.aa and (it & 1) == 0)
        n = self._outer_n()
        links = self._links
        local_pbc = any(self._spec._periodicity)
        if not links:
            dev.collide(save_macro=sync_req)
            dev.apply_pbc()
        else:
            cs, ds = self._calc_stream, self._data_stream
            if it > 0:
                # the neighbours' planes of the previous step must have landed (and, symmetrically,
                # the neighbours are done with the layers this step's push overwrites)
                for side, link in links.items():
                    if link.peer_runner is not None:
                        cs.wait_for_event(link.peer_runner._push_events[(it - 1) & 1])
                    else:
                        self._lib.call('slf_sem_wait', link.my_ready, link.my_consumed, cs.handle)
            if local_pbc:
                # periodic x/y inside a z-split: the PBC kernels need the whole step and must run
                # before the planes leave (they complete the x/y ghost ring of those planes)
                dev.collide(save_macro=sync_req, stream=cs)
                dev.apply_pbc(stream=cs)
                self._push_all(swap, cs, it)
            else:
                lo, hi = 1, n - 1
                if 0 in links:
                    dev.collide(save_macro=sync_req, z_begin=1, z_end=2, stream=cs)
                    lo = 2
                if 1 in links and n - 2 >= lo:
                    dev.collide(save_macro=sync_req, z_begin=n - 2, z_end=n - 1, stream=cs)
                    hi = n - 2
                self._slab_event.record(cs)
                ds.wait_for_event(self._slab_event)
                self._push_all(swap, ds, it)
                if hi > lo:
                    dev.collide(save_macro=sync_req, z_begin=lo, z_end=hi, stream=cs)
                cs.wait_for_event(self._push_events[it & 1])
        dev.iteration += 1
        self._sim.iteration += 1

    def _push_all(self, swap, stream, it):
        dev = self._dev
        _, dout = dev.io_dists()
        for side, link in self._links.items():
            if dev.aa:
                peer = link.peer_dist_a
            else:
                # the neighbour's output lattice of this step has the same parity as ours
                peer = link.peer_dist_b if (it & 1) == 0 else link.peer_dist_a
            self._lib.call('slf_halo_push', dev.plan, int(dout), link.face, peer, link.peer_lat_n,
                           link.peer_arr_n, swap, stream.handle)
            if link.peer_runner is None:
                self._lib.call('slf_sem_signal', link.peer_ready, stream.handle)
        self._push_events[it & 1].record(stream)

    def run_steps(self, n, save_last=False):
        """n steps without output; issued from C when the subdomain has no neighbours."""
        if self._links:
            for k in range(n):
                self.step(save_last and k == n - 1)
        else:
            self._dev.run(n, save_last=save_last)
            self._sim.iteration += n

    def _fields_to_host(self):
        self._dev.fields_to_host()

    def need_quit(self):
        """subdomain_runner.py:1604-1623."""
        if self.config.max_iters > 0:
            it = self._sim.iteration
            if it == self.config.max_iters - 1:
                self._sim.need_sync_flag = True      # request macro fields in the last step
            elif it >= self.config.max_iters:
                return True
        if self._quit_event is not None and self._quit_event.is_set():
            return True
        return False

    def main_step(self):
        """One pass of the main loop body (subdomain_runner.py:1673-1748); returns True to quit.
        As in the reference, output written at iteration N holds the macroscopic fields the
        N-th step computed from its pre-collision populations."""
        output_req = self._sim.need_output()
        sync_req, fields_req = self._sim.need_sync_fields()
        self.step(fields_req)
        if sync_req:
            self._fields_to_host()
        quit_ = self.need_quit()
        if quit_:
            self.sync()
            if output_req or sync_req:
                self._output.save(self._sim.iteration)
            return True
        if output_req:
            self.sync()
            self._output.save(self._sim.iteration)
        elif sync_req:
            self._output.mask_nonfluid_nodes()
        self._sim.after_step(self)
        return False

    def sync(self):
        self._calc_stream.synchronize()
        self._data_stream.synchronize()
