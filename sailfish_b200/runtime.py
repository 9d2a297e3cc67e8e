"""Device-side state and step scheduling for one subdomain.

``DeviceLattice`` is the core the self-contained ``SubdomainRunner`` mirror, the tests
and ``bench.py`` share: it owns the lattice/field/map buffers on the GPU and issues the
same kernel sequence as the reference's ``SubdomainRunner.step``
(sailfish/subdomain_runner.py:960-1058): CollideAndPropagate (boundary slabs first when
the subdomain has neighbours), then the per-axis PBC kernels, with the AA/AB parity
rules of lb_single.py:96-185.
"""
import ctypes

import numpy as np

from sailfish_b200 import backend_b200 as bk
from sailfish_b200 import params as P

_vp = ctypes.c_void_p


class DeviceLattice(object):
    def __init__(self, params, node_map, rho, v, device=0, backend=None, host_fields=None):
        """params: SlfParams; node_map: uint32 array (array shape, ghosts + padding
        included); rho, v: initial macroscopic fields of the same shape (host arrays)."""
        self.be = backend or bk.B200Backend(None, device)
        self.lib = bk.lib()
        self.p = params
        self.plan = self.be.build(params)
        self.n = params.dist_size
        self.dim = params.dim
        self.Q = params.Q
        self.aa = params.access_pattern == P.ACCESS_IDS['AA']
        self.shape = ((params.arr_nz,) if self.dim == 3 else ()) + (params.arr_ny, params.arr_nx)
        f32 = np.float32
        self.h_map = np.ascontiguousarray(node_map, dtype=np.uint32).reshape(self.shape)
        if host_fields is not None:
            # adopt the caller's (pinned, ghost-inclusive) arrays: no second host copy
            self.h_rho, self.h_v = host_fields[0], list(host_fields[1])
            assert self.h_rho.shape == self.shape and self.h_rho.dtype == f32
        else:
            self.h_rho = self.be.alloc_async_host_buf(self.shape, f32)
            self.h_rho[:] = np.asarray(rho, dtype=f32).reshape(self.shape)
            self.h_v = []
            for c in v:
                a = self.be.alloc_async_host_buf(self.shape, f32)
                a[:] = np.asarray(c, dtype=f32).reshape(self.shape)
                self.h_v.append(a)
        self.d_map = self.be.alloc_buf(like=self.h_map)
        self.d_rho = self.be.alloc_buf(like=self.h_rho)
        self.d_v = [self.be.alloc_buf(like=a) for a in self.h_v]
        nbytes = self.Q * self.n * 4
        self.d_dist_a = self.be.alloc_buf(size=nbytes)
        self.d_dist_b = None if self.aa else self.be.alloc_buf(size=nbytes)
        self.lib.call('slf_plan_bind_map', self.plan, int(self.d_map))
        self.be._bound_map = int(self.d_map)
        self.iteration = 0
        self.stream = self.be.make_stream()
        self.set_initial_conditions()

    # -- kernels ----------------------------------------------------------------------
    def _vz(self):
        return int(self.d_v[2]) if self.dim == 3 else 0

    def set_initial_conditions(self):
        """lb_single.py:72-94: both lattices for AB."""
        for d in (self.d_dist_a, self.d_dist_b):
            if d is None:
                continue
            self.lib.call('slf_set_initial_conditions', self.plan, int(d), int(self.d_v[0]), int(self.d_v[1]),
                          self._vz(), int(self.d_rho), int(self.d_map), self.stream.handle)
        self.iteration = 0

    def io_dists(self):
        if self.aa:
            return self.d_dist_a, self.d_dist_a
        if (self.iteration & 1) == 0:
            return self.d_dist_a, self.d_dist_b
        return self.d_dist_b, self.d_dist_a

    def current_dist_buf(self):
        if self.aa or (self.iteration & 1) == 0:
            return self.d_dist_a
        return self.d_dist_b

    def collide(self, save_macro=False, z_begin=0, z_end=-1, stream=None):
        din, dout = self.io_dists()
        self.lib.call('slf_collide_and_propagate', self.plan, int(self.d_map), int(din), int(dout),
                      int(self.d_rho), int(self.d_v[0]), int(self.d_v[1]), self._vz(),
                      1 if save_macro else 0, self.iteration, z_begin, z_end,
                      (stream or self.stream).handle)

    def apply_pbc(self, stream=None):
        """subdomain_runner.py:984-1009 + lb_single.py:153-185."""
        _, dout = self.io_dists()
        swap = int(self.aa and (self.iteration & 1) == 0)
        for axis in range(self.dim):
            if self.p.periodic[axis]:
                self.lib.call('slf_apply_periodic_boundary_conditions', self.plan, int(dout), axis, swap,
                              (stream or self.stream).handle)

    def step(self, save_macro=False):
        self.collide(save_macro)
        self.apply_pbc()
        self.iteration += 1

    def run(self, steps, save_last=True):
        """`steps` time steps issued from C (slf_run_steps: CUDA-graph replay of step pairs)."""
        if steps <= 0:
            return
        self.lib.call('slf_run_steps', self.plan, int(self.d_map), int(self.d_dist_a),
                      int(self.d_dist_b) if self.d_dist_b else 0, int(self.d_rho), int(self.d_v[0]),
                      int(self.d_v[1]), self._vz(), self.iteration, int(steps), int(bool(save_last)),
                      self.stream.handle)
        self.iteration += steps

    def run_python_loop(self, steps, save_last=True):
        for k in range(steps):
            self.step(save_macro=(save_last and k == steps - 1))

    def sync(self):
        self.stream.synchronize()

    # -- host access --------------------------------------------------------------------
    def fields_to_host(self):
        self.sync()
        self.be.from_buf(self.d_rho)
        for b in self.d_v:
            self.be.from_buf(b)
        return self.h_rho, self.h_v

    def get_dist(self):
        """Current populations as a [Q] + shape host array (natural slots for AB and for AA
        after an even number of steps; cf. _debug_get_dist, subdomain_runner.py:1363-1381)."""
        self.sync()
        out = np.zeros((self.Q,) + self.shape, dtype=np.float32)
        self.lib.call('slf_memcpy_d2h', out.ctypes.data_as(_vp), int(self.current_dist_buf()), int(out.nbytes))
        return out

    def set_dist(self, dist):
        dist = np.ascontiguousarray(dist, dtype=np.float32)
        assert dist.shape == (self.Q,) + self.shape
        self.sync()
        self.lib.call('slf_memcpy_h2d', int(self.current_dist_buf()), dist.ctypes.data_as(_vp), int(dist.nbytes))


class DeviceShanChenLattice(object):
    """Binary Shan-Chen mixture on the device: two lattices, rho / phi / v fields and the
    step of NNSubdomainRunner.step (sailfish/subdomain_runner.py:2102-2197):
    ShanChenPrepareMacroFields -> macro PBC (+ field halo) -> ShanChenCollideAndPropagate0/1 ->
    distribution PBC (+ plane halo).  With ``fused`` (default) the two collide launches are one
    kernel and the macro pass stores only the densities except on steps whose velocity is wanted."""

    fused = True

    def __init__(self, params, node_map, rho, phi, v, device=0, backend=None, host_fields=None):
        self.be = backend or bk.B200Backend(None, device)
        self.lib = bk.lib()
        self.p = params
        self.plan = self.be.build(params)
        self.n = params.dist_size
        self.dim = params.dim
        self.Q = params.Q
        self.aa = params.access_pattern == P.ACCESS_IDS['AA']
        self.shape = ((params.arr_nz,) if self.dim == 3 else ()) + (params.arr_ny, params.arr_nx)
        f32 = np.float32
        self.h_map = np.ascontiguousarray(node_map, dtype=np.uint32).reshape(self.shape)
        if host_fields is not None:
            self.h_rho, self.h_phi, self.h_v = host_fields[0], host_fields[1], list(host_fields[2])
        else:
            def pin(a):
                b = self.be.alloc_async_host_buf(self.shape, f32)
                b[:] = np.asarray(a, dtype=f32).reshape(self.shape)
                return b
            self.h_rho, self.h_phi, self.h_v = pin(rho), pin(phi), [pin(c) for c in v]
        self.d_map = self.be.alloc_buf(like=self.h_map)
        self.d_rho = self.be.alloc_buf(like=self.h_rho)
        self.d_phi = self.be.alloc_buf(like=self.h_phi)
        self.d_v = [self.be.alloc_buf(like=a) for a in self.h_v]
        nbytes = self.Q * self.n * 4
        self.d_dist = [[self.be.alloc_buf(size=nbytes), None if self.aa else self.be.alloc_buf(size=nbytes)]
                       for _ in range(2)]
        self.iteration = 0
        self.stream = self.be.make_stream()
        self.set_initial_conditions()

    def _vz(self):
        return int(self.d_v[2]) if self.dim == 3 else 0

    def set_initial_conditions(self):
        for k in (0, 1):
            if self.d_dist[0][k] is None:
                continue
            self.lib.call('slf_sc_set_initial_conditions', self.plan, int(self.d_map), int(self.d_dist[0][k]),
                          int(self.d_dist[1][k]), int(self.d_v[0]), int(self.d_v[1]), self._vz(),
                          int(self.d_rho), int(self.d_phi), self.stream.handle)
        self.iteration = 0

    def io_dists(self, g):
        a, b = self.d_dist[g]
        if b is None:
            return a, a
        return (a, b) if (self.iteration & 1) == 0 else (b, a)

    def current_dist_buf(self, g):
        a, b = self.d_dist[g]
        return a if (b is None or (self.iteration & 1) == 0) else b

    def macro(self, z_begin=0, z_end=-1, stream=None, with_v=True):
        options = 0 if (with_v or not self.fused) else P.OPTION_DENSITIES_ONLY
        self.lib.call('slf_sc_prepare_macro_fields', self.plan, int(self.d_map), int(self.io_dists(0)[0]),
                      int(self.io_dists(1)[0]), int(self.d_rho), int(self.d_phi), int(self.d_v[0]),
                      int(self.d_v[1]), self._vz(), options, self.iteration, z_begin, z_end,
                      (stream or self.stream).handle)

    def macro_pbc(self, stream=None):
        for axis in range(self.dim):
            if self.p.periodic[axis]:
                for f in (self.d_rho, self.d_phi):
                    self.lib.call('slf_apply_macro_periodic_boundary_conditions', self.plan, int(f), axis,
                                  (stream or self.stream).handle)

    def collide(self, z_begin=0, z_end=-1, stream=None):
        if self.fused:
            (i0, o0), (i1, o1) = self.io_dists(0), self.io_dists(1)
            self.lib.call('slf_sc_collide_and_propagate_fused', self.plan, int(self.d_map), int(i0), int(o0),
                          int(i1), int(o1), int(self.d_rho), int(self.d_phi), int(self.d_v[0]), int(self.d_v[1]),
                          self._vz(), 0, self.iteration, z_begin, z_end, (stream or self.stream).handle)
            return
        for g in (0, 1):
            din, dout = self.io_dists(g)
            self.lib.call('slf_sc_collide_and_propagate', self.plan, g, int(self.d_map), int(din), int(dout),
                          int(self.d_rho), int(self.d_phi), int(self.d_v[0]), int(self.d_v[1]), self._vz(), 0,
                          self.iteration, z_begin, z_end, (stream or self.stream).handle)

    def dist_pbc(self, stream=None):
        swap = int(self.aa and (self.iteration & 1) == 0)
        for axis in range(self.dim):
            if self.p.periodic[axis]:
                for g in (0, 1):
                    self.lib.call('slf_apply_periodic_boundary_conditions', self.plan, int(self.io_dists(g)[1]),
                                  axis, swap, (stream or self.stream).handle)

    def step(self, save=True):
        """``save``: store the velocity of this step (rho and phi always are)."""
        self.macro(with_v=save)
        self.macro_pbc()
        self.collide()
        self.dist_pbc()
        self.iteration += 1

    def run(self, steps, save_last=True):
        for k in range(steps):
            self.step(save_last and k == steps - 1)

    def sync(self):
        self.stream.synchronize()

    def fields_to_host(self):
        self.sync()
        for b in [self.d_rho, self.d_phi] + self.d_v:
            self.be.from_buf(b)
        return self.h_rho, self.h_phi, self.h_v

    def get_dist(self, g):
        self.sync()
        out = np.zeros((self.Q,) + self.shape, dtype=np.float32)
        self.lib.call('slf_memcpy_d2h', out.ctypes.data_as(_vp), int(self.current_dist_buf(g)), int(out.nbytes))
        return out
