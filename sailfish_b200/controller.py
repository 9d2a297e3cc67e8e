"""``LBSimulationController``: mirror of sailfish/controller.py:272-830 for one box.

Same constructor and ``run()`` contract as the reference (lb_class, lb_geo, default_config;
command-line flags contributed by every class in ``lb_class.mro()`` and by the geometry
class).  Process model: one process per GPU.  Launched plainly it drives every subdomain
from this process (round-robin over the visible GPUs -- the reference's ``gpus[i % n]``
rule, master.py:106-117); under ``torchrun`` (WORLD_SIZE > 1) rank r owns subdomain r and
the ranks exchange CUDA-IPC handles once at start-up, after which all halo traffic is
GPU-to-GPU.  The reference's cluster/PBS/execnet/ZeroMQ control plane is out of scope.
"""
import ctypes
import os
import sys
import time

import numpy as np

from sailfish_b200 import backend_b200 as bk
from sailfish_b200 import config as cfgmod
from sailfish_b200.geo import LBGeometry2D, LBGeometry3D
from sailfish_b200.subdomain_runner import checkpoint_path, HaloLink, SubdomainRunner, LBOutput

_u64 = ctypes.c_uint64


class LBSimulationController(object):
    def __init__(self, lb_class, lb_geo=None, default_config=None):
        self._lb_class = lb_class
        self.dim = lb_class.subdomain.dim
        if lb_geo is None:
            lb_geo = LBGeometry2D if self.dim == 2 else LBGeometry3D
        self._lb_geo = lb_geo
        self._parser = cfgmod.LBConfigParser()
        g = self._parser.add_group('Runtime mode settings')
        g.add_argument('--mode', type=str, choices=['batch', 'benchmark'], default='batch')
        g.add_argument('--every', type=int, default=100, help='save results every N iterations')
        g.add_argument('--from', dest='from_', type=int, default=0)
        g.add_argument('--max_iters', type=int, default=0, help='number of iterations to run (0 = run nothing)')
        g.add_argument('--output', type=str, default='', help='base name of the .npz output files')
        g.add_argument('--sync_output', action='store_true', default=False,
                       help='write output the way the reference does (synchronous copy of the padded fields, '
                            'host-side masking, savez in the main thread) instead of the device-side pack + '
                            'writer thread')
        g.add_argument('--checkpoint_file', type=str, default='', help='base name of checkpoint files')
        g.add_argument('--single_checkpoint', action='store_true', default=False)
        g.add_argument('--restore_from', type=str, default='', help='checkpoint to start from')
        g.add_argument('--norestore_time', action='store_false', dest='restore_time', default=True)
        g.add_argument('--final_checkpoint', action='store_true', default=False)
        g.add_argument('--checkpoint_every', type=int, default=0)
        g.add_argument('--checkpoint_from', type=int, default=0)
        g.add_argument('--seed', type=int, default=int(time.time()))
        g.add_argument('--gpus', type=int, nargs='+', default=[0])
        g.add_argument('--quiet', action='store_true', default=False)
        g.add_argument('--perf_stats_every', type=int, default=1000)
        g.add_argument('--debug_single_process', action='store_true', default=True)
        g.add_argument('--precision', type=str, choices=['single'], default='single')
        g.add_argument('--mem_alignment', type=int, default=32)
        g.add_argument('--block_size', type=int, default=64, help='accepted for compatibility; unused')
        g = self._parser.add_group('Simulation-specific settings')
        for base in lb_class.mro():
            if 'add_options' in base.__dict__:
                base.add_options(g, self.dim)
        g = self._parser.add_group('Geometry settings')
        lb_geo.add_options(g)
        defaults = {}
        lb_class.update_defaults(defaults)
        self._parser.set_defaults(defaults)
        if default_config is not None:
            self._parser.set_defaults(default_config)
        self.runners = []
        self._local = False

    # -- set-up --------------------------------------------------------------------------------
    def _make_specs(self):
        """Subdomain list + face connections (controller.py:155-269 of the reference: _connect_subdomains and local
        periodicity).  Two subdomains are neighbours across an axis when one ends where the other begins (modulo the
        global size on a periodic axis) and they share the WHOLE face; edge and corner neighbours need no connection
        of their own here (dimension-ordered exchange)."""
        cfg = self.config
        self.geo = self._lb_geo(cfg)
        specs = self.geo.subdomains()
        gsize = list(self.geo.gsize)
        dim = self.dim
        per = [bool(getattr(cfg, 'periodic_' + a, False)) for a in 'xyz'[:dim]]
        for a in range(dim):
            spans = set((s.location[a], s.size[a]) for s in specs)
            whole = all(sz == gsize[a] for _, sz in spans)
            if per[a] and whole:
                for s in specs:
                    s.enable_local_periodicity(a)     # one block along this axis: PBC kernels
            if whole:
                continue
            for s in specs:
                end = s.location[a] + s.size[a]
                if end == gsize[a]:
                    if not per[a]:
                        continue
                    end = 0                            # wrap-around neighbour
                for o in specs:
                    if o is s or o.location[a] != end:
                        continue
                    if all(o.location[b] == s.location[b] and o.size[b] == s.size[b] for b in range(dim) if b != a):
                        s.connect(2 * a + 1, o.id)
                        o.connect(2 * a, s.id)
                    elif all(o.location[b] < s.location[b] + s.size[b] and s.location[b] < o.location[b] + o.size[b]
                             for b in range(dim) if b != a):
                        raise NotImplementedError('subdomains %d and %d touch across %s but do not share a whole face'
                                                  % (s.id, o.id, 'xyz'[a]))
        return specs

    def _distributed(self):
        if self._local:
            return False, 0, 1, int(os.environ.get('LOCAL_RANK', '0'))
        ws = int(os.environ.get('WORLD_SIZE', '1'))
        return ws > 1, int(os.environ.get('RANK', '0')), ws, int(os.environ.get('LOCAL_RANK', '0'))

    def setup(self, args=None, local=False):
        """local=True: this process drives every subdomain of the geometry itself on its own GPU even under
        torchrun (verification runs beside a distributed job)."""
        self._local = bool(local)
        self.config = self._parser.parse(list(args or []))
        self._lb_class.modify_config(self.config)
        np.random.seed(self.config.seed)
        specs = self._make_specs()
        self.specs = specs
        dist_mode, rank, ws, local_rank = self._distributed()
        if dist_mode:
            if len(specs) != ws:
                raise ValueError('torchrun world size %d != --subdomains %d' % (ws, len(specs)))
            mine = [specs[rank]]
            devices = {specs[rank].id: local_rank}
        else:
            mine = specs
            ngpu = max(1, bk.B200Backend.devices_count())
            gpus = [g for g in self.config.gpus if g < ngpu] or [0]
            if self._local:
                gpus = [local_rank]
            devices = dict((s.id, gpus[i % len(gpus)]) for i, s in enumerate(specs))
        if len(set(devices.values())) > 1:
            raise NotImplementedError('one process drives one GPU; use torchrun for multi-GPU runs')
        self.runners = []
        for s in mine:
            sim = self._lb_class(self.config)
            runner_cls = self._lb_class.subdomain_runner or SubdomainRunner
            r = runner_cls(sim, s, output=LBOutput(self.config, s.id), device=devices[s.id])
            r.initialize()
            self.runners.append(r)
        for r in self.runners:
            if self.config.restore_from:
                r.restore_checkpoint(checkpoint_path(self.config.restore_from, r._spec.id))
        self._link(dist_mode, rank, ws)
        for r in self.runners:
            r._pre_checkpoint = self.quiesce
        # nobody may push into a neighbour's buffers before that neighbour has finished uploading and
        # initialising them
        self.sync()
        self.barrier()
        return self

    def _link(self, dist_mode, rank, ws):
        """Exchanges buffer addresses between neighbours: plain pointers inside one process, CUDA
        IPC handles (backend_cuda.py:122-126 wraps them, never uses them) across ranks."""
        by_id = dict((r._spec.id, r) for r in self.runners)
        remote = {}

        def geometry(r):
            return dict(lat=list(reversed(r._lat_size)), arr=list(reversed(r._physical_size)))
        # subdomains with do-nothing or link-bounce nodes exchange through the windowed face push (all of them, or the
        # step phases of neighbours would not match); with a locally periodic axis the plane push stays (the windowed
        # exchange has no periodic kernels beside it)
        windowed = any(r.has_boundary_sensitive_nodes() for r in self.runners if hasattr(r, 'has_boundary_sensitive_nodes'))
        periodic = any(any(r._spec._periodicity) for r in self.runners)
        if dist_mode:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group('gloo')
            mine = self.runners[0]
            be = mine.backend
            payload = dict(id=mine._spec.id, bufs=dict((k, be.ipc_handle(v) if v else None)
                                                       for k, v in mine.export_buffers().items()),
                           windowed=windowed, periodic=periodic, **geometry(mine))
            gathered = [None] * ws
            dist.all_gather_object(gathered, payload)
            windowed = any(g['windowed'] for g in gathered)
            periodic = any(g['periodic'] for g in gathered)
            needed = set(mine._spec._face_conn.values())
            for g in gathered:
                if g['id'] in needed and g['id'] != mine._spec.id:
                    remote[g['id']] = dict(lat=g['lat'], arr=g['arr'],
                                           bufs=dict((k, int(be.ipc_handle_wrap(h)) if h else 0)
                                                     for k, h in g['bufs'].items()))
            self._dist = dist
        for r in self.runners:
            r._windowed_exchange = windowed and not periodic and hasattr(r, 'has_boundary_sensitive_nodes')
        for r in self.runners:
            for face, other in sorted(r._spec._face_conn.items()):
                axis = face // 2
                if other in by_id:
                    o = by_id[other]
                    g = geometry(o)
                    r.add_link(HaloLink(face, o.export_buffers(), g['lat'][axis], g['arr'][axis], peer_runner=o,
                                        peer_lat=g['lat'], peer_arr=g['arr']))
                else:
                    info = remote[other]
                    r.add_link(HaloLink(face, info['bufs'], info['lat'][axis], info['arr'][axis],
                                        peer_sems=info['bufs']['sems'], my_sems=int(r._sems),
                                        peer_lat=info['lat'], peer_arr=info['arr']))

    # -- running -------------------------------------------------------------------------------
    def barrier(self):
        """Host-side rendezvous of all ranks (local streams drained first)."""
        if getattr(self, '_dist', None) is not None:
            self.sync()
            self._dist.barrier()

    def quiesce(self):
        """Every local stream drained and every rank here: no halo plane is in flight anywhere."""
        self.sync()
        self.barrier()

    def reinitialize(self, upload=True, reset_fields=False):
        """Restarts the run from the host fields (rho, v[, phi]) on every subdomain.  The neighbours write
        into each other's lattices, so nobody may start stepping (and pushing) before everybody has finished
        re-initialising, and nobody may re-initialise before everybody has stopped stepping: rendezvous on
        both sides.  The device semaphores need no reset (csrc/slf_halo.cuh)."""
        self.quiesce()
        for r in self.runners:
            r.reinitialize(upload, reset_fields)
        self.quiesce()

    def step_all(self, sync_req=False):
        """One time step on every local subdomain, phase by phase (a phase may end with an exchange
        the next phase of the *neighbours* depends on)."""
        for ph in self.runners[0].phases():
            for r in self.runners:
                getattr(r, ph)(sync_req)

    def run_steps(self, n, save_last=False):
        if len(self.runners) == 1:
            self.runners[0].run_steps(n, save_last=save_last)
        else:
            for k in range(n):
                self.step_all(save_last and k == n - 1)

    def sync(self):
        for r in self.runners:
            r.sync()

    def run(self, ignore_cmdline=False):
        self.setup([] if ignore_cmdline else sys.argv[1:])
        return self.run_loop()

    def run_loop(self):
        """The main loop of every local subdomain (after setup())."""
        cfg = self.config
        if cfg.max_iters <= 0:
            return self
        t0 = time.time()
        done = False
        while not done:
            for r in self.runners:
                r.begin_main_step()
            for ph in self.runners[0].phases():
                for r in self.runners:
                    getattr(r, ph)(r._fields_req)
            flags = [r.end_main_step() for r in self.runners]
            done = all(flags)
        self.sync()
        for r in self.runners:
            r.finish_output()
            if cfg.checkpoint_file and cfg.final_checkpoint and r._sim.iteration >= cfg.max_iters:
                r.save_checkpoint()
        self.barrier()
        dt = time.time() - t0
        self.loop_seconds = dt
        if not cfg.quiet:
            nodes = sum(r.num_fluid_nodes for r in self.runners)
            print('Simulation completed after %d iterations: %.2f MLUPS (local subdomains)'
                  % (cfg.max_iters, nodes * cfg.max_iters / dt * 1e-6))
        return self
