"""A stand-in for libslf_b200.so whose "device" is host memory and whose kernels are the CPU
oracle.  TEST INFRASTRUCTURE: lets the unmodified reference runner drive backend_b200's glue
(kernel-name -> C entry point mapping, argument orders, iteration patching, AA/AB buffers, PBC
kernel selection) end to end in the GPU-less build container.  Never used by the product."""
import ctypes

import numpy as np

from oracle import oracle as orc
from sailfish_b200 import params as P


class FakeLib(object):
    def __init__(self):
        self.bufs = {}          # address -> numpy byte array (keeps memory alive)
        self.plans = {}
        self.links = {}         # plan key -> (low, high)
        self.launches = 0
        self.calls = []
        self._capture = None
        self.graphs = []

    # -- helpers -------------------------------------------------------------------------------
    def _arr(self, ptr, dtype, count):
        ptr = int(ptr)
        if ptr == 0:
            return None
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (count,))

    def launch_count(self):
        return self.launches

    def call(self, name, *args):
        self.calls.append(name)
        if self._capture is not None and not name.startswith('slf_graph_'):
            self._capture.append((name, args))          # recorded, not run (stream capture)
            return None
        return getattr(self, '_' + name)(*args)

    # -- stream capture: the recorded calls are replayed in order -----------------------------------
    def _slf_graph_begin_capture(self, stream):
        self._capture = []

    def _slf_graph_end_capture(self, stream, out):
        self.graphs.append(self._capture)
        self._capture = None
        out._obj.value = len(self.graphs)

    def _slf_graph_launch(self, graph, stream):
        for name, args in self.graphs[int(graph) - 1]:
            getattr(self, '_' + name)(*args)

    def _slf_graph_destroy(self, graph):
        pass

    # -- runtime -------------------------------------------------------------------------------
    def _slf_init(self, dev):
        pass

    def _slf_device_count(self, out):
        out._obj.value = 1

    def _slf_device_info(self, name, n, mem, sms):
        if mem is not None:
            mem._obj.value = 180 << 30
        if sms is not None:
            sms._obj.value = 148

    def _slf_malloc(self, nbytes, out):
        a = np.zeros(max(int(nbytes), 8) + 64, dtype=np.uint8)
        addr = a.ctypes.data
        self.bufs[addr] = a
        out._obj.value = addr

    def _slf_free(self, ptr):
        self.bufs.pop(int(ptr), None)

    def _slf_memset(self, ptr, value, nbytes, stream):
        ctypes.memset(int(ptr), value, int(nbytes))

    def _slf_host_alloc(self, nbytes, out):
        a = np.zeros(max(int(nbytes), 8), dtype=np.uint8)
        self.bufs[a.ctypes.data] = a
        out._obj.value = a.ctypes.data

    def _slf_memcpy_h2d(self, dptr, hptr, nbytes):
        src = hptr.value if hasattr(hptr, 'value') else ctypes.cast(hptr, ctypes.c_void_p).value
        ctypes.memmove(int(dptr), src, int(nbytes))

    def _slf_memcpy_d2h(self, hptr, dptr, nbytes):
        dst = hptr.value if hasattr(hptr, 'value') else ctypes.cast(hptr, ctypes.c_void_p).value
        ctypes.memmove(dst, int(dptr), int(nbytes))

    def _slf_memcpy_h2d_async(self, dptr, hptr, nbytes, stream):
        self._slf_memcpy_h2d(dptr, hptr, nbytes)

    def _slf_memcpy_d2h_async(self, hptr, dptr, nbytes, stream):
        self._slf_memcpy_d2h(hptr, dptr, nbytes)

    def _slf_sync(self):
        pass

    def _slf_stream_create(self, out, prio):
        out._obj.value = 1

    def _slf_stream_sync(self, s):
        pass

    def _slf_stream_wait_event(self, s, e):
        pass

    def _slf_event_create(self, out, timing):
        out._obj.value = 1

    def _slf_event_record(self, e, s):
        pass

    def _slf_event_sync(self, e):
        pass

    def _slf_event_elapsed_ms(self, a, b, out):
        out._obj.value = 0.0

    # -- plan + kernels --------------------------------------------------------------------------
    def _slf_plan_create(self, params_ref, out):
        p = params_ref._obj
        key = len(self.plans) + 1
        self.plans[key] = p
        out._obj.value = key

    def _slf_plan_bind_map(self, plan, m):
        pass

    def _key(self, plan):
        return plan.value if hasattr(plan, 'value') else int(plan)

    def _p(self, plan):
        return self.plans[self._key(plan)]

    def _slf_plan_destroy(self, plan):
        self.plans.pop(self._key(plan), None)

    def _slf_plan_set_links(self, plan, lo, hi):
        self.links[self._key(plan)] = (int(bool(lo)), int(bool(hi)))

    def _slf_plan_set_face_links(self, plan, links):
        if not hasattr(self, 'face_links'):
            self.face_links = {}
        self.face_links[self._key(plan)] = int(links) & 0x3f

    def _slf_set_initial_conditions(self, plan, dist, vx, vy, vz, rho, m, stream):
        p = self._p(plan)
        n = p.dist_size
        f = np.float32
        v = [self._arr(vx, f, n), self._arr(vy, f, n)] + ([self._arr(vz, f, n)] if p.dim == 3 else [])
        orc.set_initial_conditions(p, self._arr(dist, f, n * p.Q), v, self._arr(rho, f, n), self._arr(m, np.uint32, n))
        self.launches += 1

    def _slf_collide_and_propagate(self, plan, m, din, dout, rho, vx, vy, vz, options, iteration, z0, z1, stream):
        p = self._p(plan)
        n = p.dist_size
        f = np.float32
        v = [self._arr(vx, f, n), self._arr(vy, f, n)] + ([self._arr(vz, f, n)] if p.dim == 3 else [])
        orc.collide_and_propagate(p, self._arr(m, np.uint32, n), self._arr(din, f, n * p.Q), self._arr(dout, f, n * p.Q),
                                  self._arr(rho, f, n), v, int(options), int(iteration), int(z0), int(z1))
        self.launches += 1

    def _slf_apply_periodic_boundary_conditions(self, plan, dist, axis, swap, stream):
        p = self._p(plan)
        if axis >= p.dim:
            return
        orc.apply_pbc(p, self._arr(dist, np.float32, p.dist_size * p.Q), int(axis), bool(swap),
                      self.links.get(self._key(plan), (0, 0)))
        self.launches += 1

    # -- native runner: step replay, peer-to-peer planes, output packing --------------------------------
    def _slf_run_steps(self, plan, m, da, db, rho, vx, vy, vz, iteration, nsteps, save_last, stream):
        """Same sequence as one_step() in slf_api.cu, nsteps times."""
        p = self._p(plan)
        aa = p.access_pattern == 1
        for k in range(int(nsteps)):
            it = int(iteration) + k
            din, dout = (da, da) if aa else ((da, db) if (it & 1) == 0 else (db, da))
            self._slf_collide_and_propagate(plan, m, din, dout, rho, vx, vy, vz,
                                            1 if (save_last and k == nsteps - 1) else 0, it, 0, -1, stream)
            for axis in range(p.dim):
                if p.periodic[axis]:
                    self._slf_apply_periodic_boundary_conditions(plan, dout, axis, int(aa and (it & 1) == 0), stream)

    def _outer(self, p):
        return p.arr_nz if p.dim == 3 else p.arr_ny

    def _slf_halo_push(self, plan, dist, face, peer_dist, peer_lat_n, peer_arr_n, with_swap, signal, stream):
        """tests/halo_ref.face_transfer is the host statement of this copy."""
        import halo_ref as halo
        from sailfish_b200 import lattice
        p = self._p(plan)
        plane = p.dist_size // self._outer(p)
        lat_n = p.lat_nz if p.dim == 3 else p.lat_ny
        grid = lattice.get({0: 'D2Q9', 1: 'D3Q19', 2: 'D3Q27'}[p.lattice])
        slots, src, dst = halo.face_transfer(grid, int(face), bool(with_swap), lat_n, int(peer_lat_n))
        mine = self._arr(dist, np.float32, p.dist_size * p.Q).reshape(p.Q, self._outer(p), plane)
        peer = self._arr(peer_dist, np.float32, plane * int(peer_arr_n) * p.Q).reshape(p.Q, int(peer_arr_n), plane)
        for s_ in slots:
            peer[s_, dst] = mine[s_, src]
        self.launches += 1

    def _slf_halo_push_face(self, plan, dist, face, peer_dist, peer_lat, peer_arr, with_swap, signal, stream):
        """Dimension-ordered exchange through any face (window rules: tests/halo_ref.face_windows)."""
        import halo_ref
        from sailfish_b200 import lattice
        p = self._p(plan)
        D = p.dim
        grid = lattice.get({0: 'D2Q9', 1: 'D3Q19', 2: 'D3Q27'}[p.lattice])
        lat = [p.lat_nx, p.lat_ny, p.lat_nz][:D]
        arr = [p.arr_nx, p.arr_ny, p.arr_nz][:D]
        plat = [int(peer_lat[k]) for k in range(D)]
        parr = [int(peer_arr[k]) for k in range(D)]
        mine = self._arr(dist, np.float32, p.dist_size * p.Q).reshape([p.Q] + arr[::-1])
        peer = self._arr(peer_dist, np.float32, int(np.prod(parr)) * p.Q).reshape([p.Q] + parr[::-1])
        links = getattr(self, 'face_links', {}).get(self._key(plan), 0x3f)
        for slot, src, dst in halo_ref.face_windows(grid, lat, int(face), bool(with_swap), plat, links):
            peer[(slot,) + tuple(dst[::-1])] = mine[(slot,) + tuple(src[::-1])]
        self.launches += 1

    def _slf_halo_push_field(self, plan, field, face, peer_field, peer_lat_n, stream):
        p = self._p(plan)
        plane = p.dist_size // self._outer(p)
        lat_n = p.lat_nz if p.dim == 3 else p.lat_ny
        hi = int(face) & 1
        mine = self._arr(field, np.float32, p.dist_size).reshape(self._outer(p), plane)
        peer = self._arr(peer_field, np.float32, plane * int(peer_lat_n)).reshape(int(peer_lat_n), plane)
        peer[0 if hi else int(peer_lat_n) - 1] = mine[lat_n - 2 if hi else 1]
        self.launches += 1

    def _slf_sem_signal(self, sem, stream):
        pass

    def _slf_sem_wait(self, ready_a, ready_b, seq, status, ahead, stream):
        pass

    def _slf_sem_status(self, status, err):
        err._obj.value = 0

    def _slf_pack_field(self, plan, m, field, out, stream):
        p = self._p(plan)
        shape = ((p.arr_nz,) if p.dim == 3 else ()) + (p.arr_ny, p.arr_nx)
        lat = ((p.lat_nz,) if p.dim == 3 else ()) + (p.lat_ny, p.lat_nx)
        real = tuple(slice(1, n - 1) for n in lat)
        f = self._arr(field, np.float32, p.dist_size).reshape(shape)[real]
        types = (self._arr(m, np.uint32, p.dist_size).reshape(shape)[real]) & ((1 << p.bits_type) - 1)
        wet_ids = [p.type_id[P.NODE_CLASS_NAMES.index(n)] for n in (
            '_NTFluid', 'NTHalfBBWall', 'NTEquilibriumVelocity', 'NTEquilibriumDensity', 'NTRegularizedVelocity',
            'NTRegularizedDensity', 'NTDoNothing', 'NTZouHeVelocity', 'NTZouHeDensity', 'NTCopy', 'NTYuOutflow', 'NTWallTMS')]
        dense = np.where(np.isin(types, wet_ids), f, np.float32(np.nan)).astype(np.float32)
        self._arr(out, np.float32, dense.size)[:] = dense.reshape(-1)
        self.launches += 1

    # -- ghost-layer pack / unpack (the reference's own halo path in plugin mode) ---------------------
    def _xfer(self, fn, plan, dist, face, base_gx, base_other, max_lx, max_other, buf, swap):
        p = self._p(plan)
        nd = {3: 1, 5: 3, 9: 3, 19: 5, 27: 9}[p.Q]
        d = self._arr(dist, np.float32, p.dist_size * p.Q)
        b = self._arr(buf, np.float32, nd * int(max_lx) * int(max_other))
        fn(p, d, int(face), int(base_gx), int(base_other), int(max_lx), int(max_other), b, bool(swap))
        self.launches += 1

    def _slf_collect_continuous_data(self, plan, dist, face, base_gx, base_other, max_lx, max_other, buf, swap, stream):
        self._xfer(orc.collect, plan, dist, face, base_gx, base_other, max_lx, max_other, buf, swap)

    def _slf_distribute_continuous_data(self, plan, dist, face, base_gx, base_other, max_lx, max_other, buf, swap,
                                        stream):
        self._xfer(orc.distribute, plan, dist, face, base_gx, base_other, max_lx, max_other, buf, swap)

    def _slf_collect_sparse_data(self, plan, idx, dist, buf, max_idx, stream):
        p = self._p(plan)
        n = int(max_idx)
        i = self._arr(idx, np.uint32, n)
        self._arr(buf, np.float32, n)[:] = self._arr(dist, np.float32, p.dist_size * p.Q)[i]
        self.launches += 1

    def _slf_distribute_sparse_data(self, plan, idx, dist, buf, max_idx, stream):
        p = self._p(plan)
        n = int(max_idx)
        i = self._arr(idx, np.uint32, n)
        self._arr(dist, np.float32, p.dist_size * p.Q)[i] = self._arr(buf, np.float32, n)
        self.launches += 1

    def _xfer_macro(self, fn, plan, field, face, base_gx, base_other, max_lx, max_other, buf):
        p = self._p(plan)
        mo = 1 if p.dim == 2 else int(max_other)
        fn(p, self._arr(field, np.float32, p.dist_size), int(face), int(base_gx), int(base_other), int(max_lx), mo,
           self._arr(buf, np.float32, int(max_lx) * mo))
        self.launches += 1

    def _slf_collect_continuous_macro_data(self, plan, field, face, base_gx, base_other, max_lx, max_other, buf, stream):
        self._xfer_macro(orc.collect_macro, plan, field, face, base_gx, base_other, max_lx, max_other, buf)

    def _slf_distribute_continuous_macro_data(self, plan, field, face, base_gx, base_other, max_lx, max_other, buf,
                                              stream):
        self._xfer_macro(orc.distribute_macro, plan, field, face, base_gx, base_other, max_lx, max_other, buf)

    # -- binary Shan-Chen ------------------------------------------------------------------------
    def _v(self, p, vx, vy, vz):
        f, n = np.float32, p.dist_size
        return [self._arr(vx, f, n), self._arr(vy, f, n)] + ([self._arr(vz, f, n)] if p.dim == 3 else [])

    def _slf_sc_set_initial_conditions(self, plan, m, d1, d2, vx, vy, vz, rho, phi, stream):
        p = self._p(plan)
        f, n = np.float32, p.dist_size
        orc.sc_set_initial_conditions(p, self._arr(m, np.uint32, n), self._arr(d1, f, n * p.Q),
                                      self._arr(d2, f, n * p.Q), self._v(p, vx, vy, vz),
                                      self._arr(rho, f, n), self._arr(phi, f, n))
        self.launches += 1

    def _slf_sc_prepare_macro_fields(self, plan, m, d1, d2, rho, phi, vx, vy, vz, options, iteration, z0, z1, stream):
        p = self._p(plan)
        f, n = np.float32, p.dist_size
        orc.sc_prepare_macro_fields(p, self._arr(m, np.uint32, n), self._arr(d1, f, n * p.Q),
                                    self._arr(d2, f, n * p.Q), self._arr(rho, f, n), self._arr(phi, f, n),
                                    self._v(p, vx, vy, vz), int(options), int(iteration), int(z0), int(z1))
        self.launches += 1

    def _slf_sc_collide_and_propagate(self, plan, grid_idx, m, din, dout, rho, phi, vx, vy, vz, options,
                                      iteration, z0, z1, stream):
        p = self._p(plan)
        f, n = np.float32, p.dist_size
        orc.sc_collide_and_propagate(p, int(grid_idx), self._arr(m, np.uint32, n), self._arr(din, f, n * p.Q),
                                     self._arr(dout, f, n * p.Q), self._arr(rho, f, n), self._arr(phi, f, n),
                                     self._v(p, vx, vy, vz), int(options), int(iteration), int(z0), int(z1))
        self.launches += 1

    def _slf_sc_collide_and_propagate_fused(self, plan, m, d1i, d1o, d2i, d2o, rho, phi, vx, vy, vz, options,
                                            iteration, z0, z1, stream):
        """The fused kernel recomputes the common velocity from the populations; here: the oracle's macro pass
        into scratch fields (or into vx.. when the velocity is to be saved), then the two collide passes."""
        p = self._p(plan)
        f, n = np.float32, p.dist_size
        mp = self._arr(m, np.uint32, n)
        a1, a2 = self._arr(d1i, f, n * p.Q), self._arr(d2i, f, n * p.Q)
        r, ph = self._arr(rho, f, n), self._arr(phi, f, n)
        save = int(options) & 1
        v = self._v(p, vx, vy, vz) if save else [np.zeros(n, f) for _ in range(p.dim)]
        r2, ph2 = r.copy(), ph.copy()       # the macro pass must not disturb the (halo-completed) density fields
        orc.sc_prepare_macro_fields(p, mp, a1, a2, r2, ph2, v, 0, int(iteration), int(z0), int(z1))
        orc.sc_collide_and_propagate(p, 0, mp, a1, self._arr(d1o, f, n * p.Q), r, ph, v, 0, int(iteration), int(z0), int(z1))
        orc.sc_collide_and_propagate(p, 1, mp, a2, self._arr(d2o, f, n * p.Q), r, ph, v, 0, int(iteration), int(z0), int(z1))
        self.launches += 1

    def _slf_apply_macro_periodic_boundary_conditions(self, plan, field, axis, stream):
        p = self._p(plan)
        if axis >= p.dim:
            return
        orc.apply_macro_pbc(p, self._arr(field, np.float32, p.dist_size), int(axis),
                            self.links.get(self._key(plan), (0, 0)))
        self.launches += 1


def install_fake(monkeypatch=None):
    from sailfish_b200 import backend_b200 as bk
    fake = FakeLib()
    if monkeypatch is not None:
        monkeypatch.setattr(bk, '_LIB', fake)
    else:
        bk._LIB = fake
    return fake
