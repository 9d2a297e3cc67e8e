"""A stand-in for libslf_b200.so whose "device" is host memory and whose kernels are the CPU
oracle.  TEST INFRASTRUCTURE: lets the unmodified reference runner drive backend_b200's glue
(kernel-name -> C entry point mapping, argument orders, iteration patching, AA/AB buffers, PBC
kernel selection) end to end in the GPU-less build container.  Never used by the product."""
import ctypes

import numpy as np

from oracle import oracle as orc


class FakeLib(object):
    def __init__(self):
        self.bufs = {}          # address -> numpy byte array (keeps memory alive)
        self.plans = {}
        self.launches = 0
        self.calls = []

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
        return getattr(self, '_' + name)(*args)

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

    def _p(self, plan):
        return self.plans[plan.value if hasattr(plan, 'value') else int(plan)]

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
        orc.apply_pbc(p, self._arr(dist, np.float32, p.dist_size * p.Q), int(axis), bool(swap))
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

    def _slf_apply_macro_periodic_boundary_conditions(self, plan, field, axis, stream):
        p = self._p(plan)
        if axis >= p.dim:
            return
        orc.apply_macro_pbc(p, self._arr(field, np.float32, p.dist_size), int(axis))
        self.launches += 1


def install_fake(monkeypatch=None):
    from sailfish_b200 import backend_b200 as bk
    fake = FakeLib()
    if monkeypatch is not None:
        monkeypatch.setattr(bk, '_LIB', fake)
    else:
        bk._LIB = fake
    return fake
