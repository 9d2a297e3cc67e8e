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

import numpy as np

from sailfish_b200 import backend_b200 as bk
from sailfish_b200 import params as P


def checkpoint_filename(base, digits, subdomain_id, it):
    """sailfish/io.py:190-191 (numpy appends '.npz')."""
    return ('{0}.{1:0' + str(digits) + 'd}.{2}.cpoint').format(base, it, subdomain_id)


def checkpoint_path(base, subdomain_id):
    """sailfish/io.py:193-202: '<base>.last' picks a file by globbing, otherwise the name is literal."""
    import glob
    if base.endswith('.last'):
        files = sorted(glob.glob('{0}.*.{1}.cpoint.npz'.format(base[:-5], subdomain_id)))
        if not files:
            raise IOError('no checkpoint matches %s' % base)
        return files[0]
    return '{0}.{1}.cpoint.npz'.format(base, subdomain_id)

_u64 = ctypes.c_uint64


class HaloLink(object):
    """One directed neighbour relation across the split axis.

    ``peer_bufs`` maps buffer names ('dist0a', 'dist0b', [dist1a, dist1b, rho, phi]) to the
    neighbour's device pointers as seen from this process.  Same process (several subdomains
    driven by one process, the reference's "several subdomains per GPU" regtest mode):
    ``peer_runner`` is set and ordering uses CUDA events.  Other rank: the pointers are CUDA-IPC
    mappings and ordering uses the stream-ordered device semaphores slf_sem_signal/slf_sem_wait:
    ``peer_sems`` is the neighbour's semaphore block, ``my_sems`` this runner's."""

    def __init__(self, face, peer_bufs, peer_lat_n, peer_arr_n, peer_sems=0, my_sems=0, peer_runner=None,
                 peer_lat=None, peer_arr=None):
        self.face = face
        self.side = face & 1
        self.axis = face >> 1
        self.peer_bufs = dict((k, int(v) if v else 0) for k, v in peer_bufs.items())
        self.peer_lat_n = int(peer_lat_n)
        self.peer_arr_n = int(peer_arr_n)
        # the neighbour's lattice / array extents (x, y[, z]): needed for faces other than the outermost axis
        self.peer_lat = [int(v) for v in peer_lat] if peer_lat is not None else None
        self.peer_arr = [int(v) for v in peer_arr] if peer_arr is not None else None
        self.peer_sems = int(peer_sems)
        self.my_sems = int(my_sems)
        self.peer_runner = peer_runner

    # semaphore block layout (uint32 words, csrc/slf_halo.cuh): channel c in {0: populations, 1: macro fields}
    #   ready[c][face] at 8*c + face, seq[c][axis] at 16 + 4*c + axis; status (error, abort) at 32, 33
    def my_ready(self, channel):
        return self.my_sems + 4 * (8 * channel + self.face)

    def peer_ready(self, channel):
        # the neighbour receives through its opposite face
        return self.peer_sems + 4 * (8 * channel + (self.face ^ 1))


SEM_BYTES = 256
SEM_STATUS_WORD = 32


def sem_seq_word(channel, axis):
    return 16 + 4 * channel + axis


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
        self._links = {}               # side -> link across the outermost axis
        self._xlinks = {}              # face -> link (every linked face)
        self._general = False          # links across x / y faces: dimension-ordered exchange
        self._graphs = {}              # step parity -> captured pair of steps (linked subdomains)
        self._pre_checkpoint = None    # set by the controller: quiesce all subdomains / ranks
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

    def update_context(self, ctx):
        """subdomain_runner.py:182-248 of the reference: model, encoder, then the array geometry of this
        subdomain (lattice size with ghosts, padded array size, which axes wrap locally)."""
        self._sim.update_context(ctx)
        sd = self._subdomain
        sd.update_context(ctx)
        lat = list(reversed(sd.lat_shape)) + ([1] if self.dim == 2 else [])
        arr = list(reversed(sd.arr_shape)) + ([1] if self.dim == 2 else [])
        for a, n, m in zip('xyz', lat, arr):
            ctx['lat_n' + a] = n
            ctx['arr_n' + a] = m
        ctx['block_periodicity'] = list(self._spec._periodicity) + [False] * (3 - self.dim)
        ctx['sim'] = self._sim
        ctx['block'] = self._spec

    def make_params(self):
        """SlfParams of this subdomain, through the same context -> parameters function that serves the
        unmodified reference (params.from_context)."""
        ctx = {}
        self.update_context(ctx)
        return P.from_context(ctx)

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
        self._init_sync_objects()

    def _init_sync_objects(self):
        self._sems = self.backend.alloc_buf(size=SEM_BYTES)     # see HaloLink for the layout
        self._lib.call('slf_memset', int(self._sems), 0, SEM_BYTES, 0)
        self._lib.call('slf_sync')
        self._push_events = [[self.backend.make_event(self._data_stream) for _ in range(2)] for _ in range(2)]
        # block decompositions (links across x / y faces too): one event per axis and step parity
        self._axis_events = [[self.backend.make_event(self._calc_stream) for _ in range(2)] for _ in range(3)]
        self._fork_event = self.backend.make_event(self._calc_stream)

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

    def _debug_get_dist(self, output=True, grid_num=0):
        """[Q, (nz+2), ny+2, arr_nx] raw copy of one lattice buffer (subdomain_runner.py:1363-1381):
        ``output`` selects the buffer written by the last step."""
        idx = self._sim.iteration & 1
        if not output:
            idx = 1 - idx
        self.sync()
        out = np.zeros([self._sim.grid.Q] + list(self._physical_size), dtype=np.float32)
        self._lib.call('slf_memcpy_d2h', out.ctypes.data_as(bk._vp), int(self.gpu_dist(grid_num, idx)),
                       int(out.nbytes))
        return out

    def _debug_set_dist(self, dist, output=True, grid_num=0):
        """subdomain_runner.py:1383-1388."""
        idx = self._sim.iteration & 1
        if not output:
            idx = 1 - idx
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        if list(dist.shape) != [self._sim.grid.Q] + list(self._physical_size):
            raise ValueError('checkpoint lattice has shape %s, this subdomain needs %s'
                             % (dist.shape, [self._sim.grid.Q] + list(self._physical_size)))
        self.sync()
        self._lib.call('slf_memcpy_h2d', int(self.gpu_dist(grid_num, idx)), dist.ctypes.data_as(bk._vp),
                       int(dist.nbytes))

    # -- checkpoints (subdomain_runner.py:1414-1449) -----------------------------------------------
    def save_checkpoint(self):
        import pickle
        cfg = self.config
        # A neighbour's push of the step just finished lands in THIS lattice from the neighbour's stream
        # (or rank); it is only awaited at the start of the next step.  Quiesce every subdomain and rank
        # first, otherwise the file may miss one step of boundary data (all subdomains checkpoint at the
        # same iterations, so the rendezvous is collective-safe).
        if self._xlinks and self._pre_checkpoint is not None:
            self._pre_checkpoint()
        if getattr(cfg, 'single_checkpoint', False):
            fname = checkpoint_filename(cfg.checkpoint_file, 1, self._spec.id, 0)
        else:
            digits = str(int(np.log10(cfg.max_iters)) + 1) if cfg.max_iters else '7'
            fname = checkpoint_filename(cfg.checkpoint_file, digits, self._spec.id, self._sim.iteration)
        data = {'state': np.frombuffer(pickle.dumps(self._sim.get_state(), 2), dtype=np.uint8)}
        for i in range(len(self._sim.grids)):
            data['dist%da' % i] = self._debug_get_dist(True, i)
            if cfg.access_pattern == 'AB':
                data['dist%db' % i] = self._debug_get_dist(False, i)
        np.savez(fname, **data)
        return fname + '.npz'

    def restore_checkpoint(self, fname):
        import pickle
        cp = np.load(fname)
        self._sim.set_state(pickle.loads(cp['state'].tobytes()))
        if not getattr(self.config, 'restore_time', True):
            self._sim.iteration = 0
        self._dev.iteration = self._sim.iteration
        for k in cp.files:
            if k.startswith('dist'):
                self._debug_set_dist(cp[k], k.endswith('a'), int(k[4:-1]))
        self.sync()

    @property
    def num_fluid_nodes(self):
        return self._subdomain.num_fluid_nodes

    # -- halo ------------------------------------------------------------------------------------
    def export_buffers(self):
        """Device buffers neighbours write into (name -> pointer)."""
        d = self._dev
        return {'dist0a': int(d.d_dist_a), 'dist0b': int(d.d_dist_b) if d.d_dist_b else 0, 'sems': int(self._sems)}

    # Node classes that keep what they find in the slots no node streams into (boundary.mako:653-668, :862-877): a
    # whole-plane push across a face that cuts the domain boundary would overwrite those slots with the neighbour's
    # never-written ghost entries.  Subdomains holding such nodes exchange through the windowed face push instead.
    BOUNDARY_SENSITIVE = ('NTDoNothing', 'NTHalfBBWall', 'NTWallTMS')

    def has_boundary_sensitive_nodes(self):
        return bool(set(self._subdomain.type_ids_by_name()) & set(self.BOUNDARY_SENSITIVE))

    def add_link(self, link):
        outer = self.dim - 1
        self._xlinks[link.face] = link
        if link.axis == outer:
            self._links[link.face & 1] = link
        if link.axis != outer or getattr(self, '_windowed_exchange', False):
            self._general = True
        if self._general and any(self._spec._periodicity):
            raise P.UnsupportedConfiguration('native runner: a locally periodic axis next to subdomain links across x / y '
                                             'faces is not supported (split the periodic axis too, or split along %s only)'
                                             % 'xyz'[outer])
        self._drop_graphs()
        self._lib.call('slf_plan_set_links', self._dev.plan, int(0 in self._links), int(1 in self._links))
        if self._general:
            self._lib.call('slf_plan_set_face_links', self._dev.plan, sum(1 << f for f in self._xlinks))

    def _outer_n(self):
        return self._lat_size[0]

    def _uses_semaphores(self):
        return bool(self._xlinks) and all(l.peer_runner is None for l in self._xlinks.values())

    def _wait_incoming(self, channel, it_prev, stream, axis=None, ahead=None):
        """Everything the neighbours across `axis` (default: the outermost) pushed in step `it_prev` has landed
        (events inside one process, device semaphores across ranks: ONE wait kernel covers both sides)."""
        axis = self.dim - 1 if axis is None else axis
        ready = [0, 0]
        for side in (0, 1):
            link = self._xlinks.get(2 * axis + side)
            if link is None:
                continue
            if link.peer_runner is not None:
                ev = (link.peer_runner._push_events[channel][it_prev & 1] if axis == self.dim - 1 and not self._general
                      else link.peer_runner._axis_events[axis][it_prev & 1])
                stream.wait_for_event(ev)
            else:
                ready[side] = link.my_ready(channel)
        if ready[0] or ready[1]:
            sems = int(self._sems)
            # ahead = 0: pushes of the previous steps (populations across the last axis); 1: of this step too (fields,
            # and the earlier axes of a dimension-ordered exchange)
            ahead = int(channel == 1) if ahead is None else int(ahead)
            self._lib.call('slf_sem_wait', ready[0], ready[1], sems + 4 * sem_seq_word(channel, axis),
                           sems + 4 * SEM_STATUS_WORD, ahead, stream.handle)

    def check_peers(self):
        """Raises if a semaphore wait gave up (dead or diverged neighbour) since the last check."""
        if self._uses_semaphores():
            err = ctypes.c_uint32(0)
            self._lib.call('slf_sem_status', int(self._sems) + 4 * SEM_STATUS_WORD, ctypes.byref(err))
            if err.value:
                raise bk.B200Error('subdomain %d: halo wait %s -- a neighbouring rank stopped pushing ghost layers'
                                   % (self._spec.id, 'timed out' if err.value == 1 else 'was aborted'))

    def abort_waits(self):
        if self._uses_semaphores():
            self._lib.call('slf_sem_abort', int(self._sems) + 4 * SEM_STATUS_WORD)

    def reset_host_fields(self):
        """The host fields as they were before the first step: inf in the envelope, zero inside, then the
        subdomain's initial_conditions() (subdomain_runner.py:253-320, subdomain.py:455-470)."""
        views = list(self._scalar_fields) + [c for v in self._vector_fields for c in v]
        for view in views:
            base = self.field_base(view)
            if base.dtype == self.float:
                base[:] = np.inf
            view[:] = 0
        self._subdomain.init_fields(self._sim)

    def reinitialize(self, upload=True, reset_fields=False):
        """Starts the run again from the host fields (as the caller left them, or reset to the initial
        conditions): upload, SetInitialConditions, iteration 0.  With neighbours this must be bracketed by
        rendezvous of all subdomains -- use LBSimulationController.reinitialize()."""
        dev = self._dev
        if reset_fields:
            self.reset_host_fields()
        if upload:
            for b in [dev.d_rho] + list(dev.d_v) + ([dev.d_phi] if hasattr(dev, 'd_phi') else []):
                self.backend.to_buf_async(b, self._calc_stream)
        dev.set_initial_conditions()
        self._sim.iteration = 0

    # -- stepping ---------------------------------------------------------------------------------
    def phases(self):
        """What LBSimulationController issues for every local subdomain in turn.  Block decompositions exchange
        dimension by dimension (csrc/slf_halo.cuh): a process that drives several subdomains must have pushed
        axis k on all of them before any of them forwards that data along axis k + 1."""
        if not self._general:
            return ['step']
        return ['_gstep_collide'] + ['_gstep_push_%d' % k for k in range(len(self._linked_axes()))] + ['_gstep_end']

    def _linked_axes(self):
        return sorted(set(f >> 1 for f in self._xlinks))

    def _gstep_collide(self, sync_req):
        dev = self._dev
        axes = self._linked_axes()
        # the planes pushed across the LAST axis in the previous step (the earlier axes were awaited in that step)
        self._wait_incoming(0, dev.iteration - 1, self._calc_stream, axis=axes[-1], ahead=0)
        dev.collide(save_macro=sync_req, stream=self._calc_stream)

    def _gstep_push(self, k):
        dev = self._dev
        it = dev.iteration
        cs = self._calc_stream
        axes = self._linked_axes()
        axis = axes[k]
        swap = int(dev.aa and (it & 1) == 0)
        if k > 0:
            self._wait_incoming(0, it, cs, axis=axes[k - 1], ahead=1)      # forwarded data must have arrived
        _, dout = dev.io_dists()
        D = self.dim
        for side in (0, 1):
            link = self._xlinks.get(2 * axis + side)
            if link is None:
                continue
            peer = link.peer_bufs['dist0a'] if (dev.aa or (it & 1)) else link.peer_bufs['dist0b']
            plat = (ctypes.c_int * 3)(*(link.peer_lat + [1] * (3 - D)))
            parr = (ctypes.c_int * 3)(*(link.peer_arr + [1] * (3 - D)))
            self._lib.call('slf_halo_push_face', dev.plan, int(dout), link.face, peer, plat, parr, swap,
                           0 if link.peer_runner is not None else link.peer_ready(0), cs.handle)
        self._axis_events[axis][it & 1].record(cs)

    def _gstep_push_0(self, sync_req):
        self._gstep_push(0)

    def _gstep_push_1(self, sync_req):
        self._gstep_push(1)

    def _gstep_push_2(self, sync_req):
        self._gstep_push(2)

    def _gstep_end(self, sync_req):
        self._dev.iteration += 1
        self._sim.iteration += 1

    def step(self, sync_req):
        """One time step.  Cf. subdomain_runner.py:960-1058 (boundary blocks -> send -> bulk -> receive).

        With neighbours across the split axis the two layers next to them (slabs) and the interior run
        CONCURRENTLY on two streams: in the AA pattern a node owns the Q addresses it reads and writes (and
        in AB the written slots of different nodes are disjoint), incoming planes are only ever read by
        slab nodes, and outgoing planes are only ever written by slab nodes.  So

            data stream (high priority):  wait for the neighbours' planes of step t-1 -> slabs ->
                                          push planes + signal (one launch per face)
            calc stream                :  interior

        fork at the start of the step, join at its end; nothing but the join is on the critical path."""
        dev = self._dev
        it = dev.iteration
        swap = int(dev.aa and (it & 1) == 0)
        n = self._outer_n()
        links = self._links
        local_pbc = any(self._spec._periodicity)
        if self._general:
            for ph in self.phases():
                getattr(self, ph)(sync_req)
            return
        if not links:
            dev.collide(save_macro=sync_req)
            dev.apply_pbc()
        else:
            cs, ds = self._calc_stream, self._data_stream
            if local_pbc:
                # periodic x/y inside a z-split: the PBC kernels need the whole step and must run
                # before the planes leave (they complete the x/y ghost ring of those planes)
                self._wait_incoming(0, it - 1, cs)
                dev.collide(save_macro=sync_req, stream=cs)
                dev.apply_pbc(stream=cs)
                self._push_all(swap, cs, it)
            elif os.environ.get('SLF_STEP_SCHEME', 'concurrent') == 'serial':
                # slabs ahead of the interior on the calc stream, only the pushes overlap (A/B switch for measurements)
                self._wait_incoming(0, it - 1, cs)
                lo, hi = 1, n - 1
                if 0 in links:
                    dev.collide(save_macro=sync_req, z_begin=1, z_end=2, stream=cs)
                    lo = 2
                if 1 in links and n - 2 >= lo:
                    dev.collide(save_macro=sync_req, z_begin=n - 2, z_end=n - 1, stream=cs)
                    hi = n - 2
                self._fork_event.record(cs)
                ds.wait_for_event(self._fork_event)
                self._push_all(swap, ds, it)
                if hi > lo:
                    dev.collide(save_macro=sync_req, z_begin=lo, z_end=hi, stream=cs)
                cs.wait_for_event(self._push_events[0][it & 1])
            else:
                self._fork_event.record(cs)
                ds.wait_for_event(self._fork_event)
                self._wait_incoming(0, it - 1, ds)
                lo, hi = 1, n - 1
                if 0 in links:
                    dev.collide(save_macro=sync_req, z_begin=1, z_end=2, stream=ds)
                    lo = 2
                if 1 in links and n - 2 >= lo:
                    dev.collide(save_macro=sync_req, z_begin=n - 2, z_end=n - 1, stream=ds)
                    hi = n - 2
                self._push_all(swap, ds, it)
                if hi > lo:
                    dev.collide(save_macro=sync_req, z_begin=lo, z_end=hi, stream=cs)
                cs.wait_for_event(self._push_events[0][it & 1])
        dev.iteration += 1
        self._sim.iteration += 1

    def _push_all(self, swap, stream, it):
        dev = self._dev
        _, dout = dev.io_dists()
        for side, link in self._links.items():
            if dev.aa:
                peer = link.peer_bufs['dist0a']
            else:
                # the neighbour's output lattice of this step has the same parity as ours
                peer = link.peer_bufs['dist0b'] if (it & 1) == 0 else link.peer_bufs['dist0a']
            self._lib.call('slf_halo_push', dev.plan, int(dout), link.face, peer, link.peer_lat_n,
                           link.peer_arr_n, swap, 0 if link.peer_runner is not None else link.peer_ready(0),
                           stream.handle)
        self._push_events[0][it & 1].record(stream)

    # -- step replay --------------------------------------------------------------------------------
    def _drop_graphs(self):
        for g in self._graphs.values():
            self._lib.call('slf_graph_destroy', g)
        self._graphs = {}

    def _step_graph(self, parity):
        """Two steps (one of each parity) recorded from step() itself: launches, fork/join events and the
        semaphore kernels (whose targets live in device memory) depend on the iteration only through its
        parity."""
        g = self._graphs.get(parity)
        if g is None:
            dev, sim = self._dev, self._sim
            it0, sit0 = dev.iteration, sim.iteration
            self._lib.call('slf_graph_begin_capture', self._calc_stream.handle)
            try:
                self.step(False)
                self.step(False)
            finally:
                h = _u64()
                try:
                    self._lib.call('slf_graph_end_capture', self._calc_stream.handle, ctypes.byref(h))
                finally:
                    dev.iteration, sim.iteration = it0, sit0
            g = self._graphs[parity] = h.value
        return g

    def run_steps(self, n, save_last=False):
        """n steps without output: replayed CUDA graphs of step pairs -- issued from C when the subdomain
        has no neighbours, recorded from step() when its neighbours live in other ranks.  (Several
        subdomains driven by one process order their pushes with events of the other runners and are
        stepped one by one: LBSimulationController.run_steps.)"""
        if not self._xlinks:
            self._dev.run(n, save_last=save_last)
            self._sim.iteration += n
            return
        plain = n - (1 if save_last else 0)
        done = 0
        if self._uses_semaphores() and plain >= 4 and not getattr(self.config, 'no_step_graphs', False) and not os.environ.get('SLF_NO_STEP_GRAPHS'):
            g = self._step_graph(self._dev.iteration & 1)
            for _ in range(plain // 2):
                self._lib.call('slf_graph_launch', g, self._calc_stream.handle)
                self._dev.iteration += 2
                self._sim.iteration += 2
            done = 2 * (plain // 2)
        for k in range(done, n):
            self.step(save_last and k == n - 1)

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
        self.begin_main_step()
        self.step(self._fields_req)
        return self.end_main_step()

    def begin_main_step(self):
        self._output_req = self._sim.need_output()
        self._host_sync_req = bool(self._sim.need_sync_flag)     # a hook wants the fields on the host
        self._sync_req, self._fields_req = self._sim.need_sync_fields()

    def _fast_output(self):
        """The device-side pack + asynchronous writer (sailfish_b200/io.py) replaces the host round trip
        unless the simulation asked for the fields on the host (need_sync_flag) or --sync_output."""
        if getattr(self.config, 'sync_output', False) or not self._output.basename:
            return None
        if getattr(self, '_fast', None) is None:
            from sailfish_b200.io import FastOutput
            self._fast = FastOutput(self, self._output)
        return self._fast

    def finish_output(self):
        if getattr(self, '_fast', None) is not None:
            self._fast.close()
            self._fast = None

    def end_main_step(self):
        output_req, sync_req = self._output_req, self._sync_req
        if output_req and not self._host_sync_req and self._fast_output() is not None:
            self._fast.save(self._sim.iteration)
            quit_ = self.need_quit()
            if quit_:
                self.sync()
                self.finish_output()
                return True
            self._sim.after_step(self)
            self._maybe_checkpoint()
            return False
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
        self._maybe_checkpoint()
        return False

    def _maybe_checkpoint(self):
        if getattr(self.config, 'checkpoint_file', '') and self._sim.need_checkpoint():
            self.save_checkpoint()

    def sync(self):
        self._calc_stream.synchronize()
        self._data_stream.synchronize()
        self.check_peers()


class NNSubdomainRunner(SubdomainRunner):
    """Runner for models whose collision needs neighbour values of macroscopic fields (binary
    Shan-Chen, BASELINE config 5): mirror of sailfish/subdomain_runner.py:1840-2197.  Per step:
    ShanChenPrepareMacroFields -> macro PBC -> field ghost layers to the neighbours ->
    ShanChenCollideAndPropagate0/1 -> distribution PBC -> population planes to the neighbours."""

    def add_link(self, link):
        if link.axis != self.dim - 1:
            raise P.UnsupportedConfiguration('binary Shan-Chen runner: subdomains are connected along the outermost axis (%s)'
                                             % 'xyz'[self.dim - 1])
        SubdomainRunner.add_link(self, link)

    def _init_fields(self):
        sim = self._sim
        sim.rho, _ = self.make_scalar_field(name='rho')
        sim.phi, _ = self.make_scalar_field(name='phi')
        sim.v = self.make_vector_field(name='v')
        for i, c in enumerate('xyz'[:self.dim]):
            setattr(sim, 'v' + c, sim.v[i])
        self._subdomain.init_fields(sim)

    def make_params(self):
        unsupported = set(self._subdomain.type_ids_by_name()) - {'_NTFluid', '_NTGhost', '_NTUnused',
                                                                 '_NTPropagationOnly', 'NTFullBBWall'}
        if unsupported:
            raise P.UnsupportedConfiguration('Shan-Chen path supports fluid and full-way bounce-back nodes; got %s'
                                             % sorted(unsupported))
        return super(NNSubdomainRunner, self).make_params()

    def _init_compute(self):
        from sailfish_b200.runtime import DeviceShanChenLattice
        self._params = self.make_params()
        sim = self._sim
        fb = self.field_base
        self._dev = DeviceShanChenLattice(self._params, self._subdomain.encoded_map(), fb(sim.rho), fb(sim.phi),
                                          [fb(c) for c in sim.v], backend=self.backend,
                                          host_fields=(fb(sim.rho), fb(sim.phi), [fb(c) for c in sim.v]))
        self._calc_stream = self._dev.stream
        self._data_stream = self.backend.make_stream(high_priority=True)
        self._init_sync_objects()

    def gpu_field(self, field):
        if isinstance(field, list):
            return list(self._dev.d_v)
        return self._dev.d_phi if field is self._sim.phi else self._dev.d_rho

    def gpu_dist(self, num, copy):
        a, b = self._dev.d_dist[num]
        return a if (copy == 0 or b is None) else b

    def export_buffers(self):
        d = self._dev
        out = {'rho': int(d.d_rho), 'phi': int(d.d_phi), 'sems': int(self._sems)}
        for g in (0, 1):
            out['dist%da' % g] = int(d.d_dist[g][0])
            out['dist%db' % g] = int(d.d_dist[g][1]) if d.d_dist[g][1] else 0
        return out

    def phases(self):
        """A step has a mid-step exchange (macro fields), so a process that drives several
        subdomains must finish phase 1 on all of them before phase 2 (controller.step_all)."""
        return ['step_macro', 'step_collide']

    def step(self, sync_req):
        self.step_macro(sync_req)
        self.step_collide(sync_req)

    def _ranges(self):
        """(interior, [slabs]) as layer ranges along the split axis: the slabs are the real layers next to a linked
        face -- the only nodes whose inputs arrive from a neighbour."""
        n = self._outer_n()
        links = self._links
        lo, hi = (2 if 0 in links else 0), (n - 2 if 1 in links else n)
        slabs = []
        if 0 in links:
            slabs.append((1, 2))
        if 1 in links and n - 2 >= (2 if 0 in links else 1):
            slabs.append((n - 2, n - 1))
        else:
            hi = n if 1 not in links else max(hi, lo)
        return (lo, max(lo, hi)), slabs

    def step_macro(self, sync_req):
        """Densities (and velocity on output steps) of this step, then the field ghost layers to the neighbours.
        The interior goes FIRST: it needs nothing from the neighbours, so their population planes of the previous
        step -- pushed at the very end of it -- travel while it runs; only the slabs wait for them."""
        dev = self._dev
        it = dev.iteration
        cs = self._calc_stream
        links = self._links
        if not links:
            dev.macro(stream=cs, with_v=bool(sync_req))
            dev.macro_pbc(stream=cs)
            return
        (lo, hi), slabs = self._ranges()
        if hi > lo:
            dev.macro(z_begin=lo, z_end=hi, stream=cs, with_v=bool(sync_req))
        self._wait_incoming(0, it - 1, cs)
        for z0, z1 in slabs:
            dev.macro(z_begin=z0, z_end=z1, stream=cs, with_v=bool(sync_req))
        dev.macro_pbc(stream=cs)
        for side, link in links.items():
            for name, buf in (('rho', dev.d_rho), ('phi', dev.d_phi)):
                self._lib.call('slf_halo_push_field', dev.plan, int(buf), link.face, link.peer_bufs[name],
                               link.peer_lat_n, cs.handle)
            if link.peer_runner is None:
                self._lib.call('slf_sem_signal', link.peer_ready(1), cs.handle)
        self._push_events[1][it & 1].record(cs)

    def step_collide(self, sync_req):
        """Relaxation + streaming of both components.  Again the interior first (its psi stencil reaches the slabs'
        densities, which are local, but not the neighbours' ghost layers), then the slabs once those have landed."""
        dev = self._dev
        it = dev.iteration
        cs = self._calc_stream
        links = self._links
        if links:
            (lo, hi), slabs = self._ranges()
            if hi > lo:
                dev.collide(z_begin=lo, z_end=hi, stream=cs)
            self._wait_incoming(1, it, cs)
            for z0, z1 in slabs:
                dev.collide(z_begin=z0, z_end=z1, stream=cs)
        else:
            dev.collide(stream=cs)
        dev.dist_pbc(stream=cs)
        if links:
            swap = int(dev.aa and (it & 1) == 0)
            for side, link in links.items():
                for g in (0, 1):
                    _, dout = dev.io_dists(g)
                    if dev.aa:
                        peer = link.peer_bufs['dist%da' % g]
                    else:
                        peer = link.peer_bufs['dist%db' % g] if (it & 1) == 0 else link.peer_bufs['dist%da' % g]
                    # the second lattice's copy carries the signal: both have landed when it fires
                    sig = link.peer_ready(0) if (g == 1 and link.peer_runner is None) else 0
                    self._lib.call('slf_halo_push', dev.plan, int(dout), link.face, peer, link.peer_lat_n,
                                   link.peer_arr_n, swap, sig, cs.handle)
            self._push_events[0][it & 1].record(cs)
        dev.iteration += 1
        self._sim.iteration += 1

    def run_steps(self, n, save_last=False):
        plain = n - (1 if save_last else 0)
        done = 0
        if ((not self._links or self._uses_semaphores()) and plain >= 4
                and not getattr(self.config, 'no_step_graphs', False) and not os.environ.get('SLF_NO_STEP_GRAPHS')):
            g = self._step_graph(self._dev.iteration & 1)
            for _ in range(plain // 2):
                self._lib.call('slf_graph_launch', g, self._calc_stream.handle)
                self._dev.iteration += 2
                self._sim.iteration += 2
            done = 2 * (plain // 2)
        for k in range(done, n):
            self.step(save_last and k == n - 1)
