"""ctypes wrapper of the CPU oracle (oracle/lbm_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compiles liboracle_f32.so / liboracle_f64.so with gcc (Makefile in this dir)."""
    args = ['make', '-C', _HERE, '-s']
    if force:
        args.append('-B')
    subprocess.check_call(args)


def _lib(dtype):
    dtype = np.dtype(dtype)
    key = 'f32' if dtype == np.float32 else 'f64'
    if key not in _LIBS:
        path = os.path.join(_HERE, 'liboracle_%s.so' % key)
        if not os.path.exists(path):
            build()
        _LIBS[key] = ctypes.CDLL(path)
        assert _LIBS[key].orc_sizeof_real() == dtype.itemsize
    return _LIBS[key]


# the CPUs this process may use, read before any thread is pinned (pinning the main thread shrinks
# what sched_getaffinity reports afterwards)
HOST_CPUS = sorted(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else list(range(os.cpu_count() or 1))


def omp_config(nthreads=0, pin=True, dtype=np.float32, cpus=None):
    """Sets the OpenMP thread count of the oracle library (0 = leave) and the threads' CPU affinity -- pin:
    thread k on the k-th CPU of `cpus` (default: every CPU the process may use); not pin: free to move within
    `cpus`.  Returns the number of threads a parallel region really gets."""
    cpus = list(cpus if cpus is not None else HOST_CPUS)
    arr = (ctypes.c_int * len(cpus))(*cpus)
    return int(_lib(dtype).orc_omp_config(int(nthreads), arr, len(cpus), int(bool(pin))))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError('oracle %s failed with code %d' % (what, rc))


def collide_and_propagate(params, node_map, dist_in, dist_out, rho, v, options, iteration,
                          z_begin=0, z_end=-1):
    """One CollideAndPropagate launch (lb_single_fluid.mako:161-229).  For the AA
    pattern pass the same array as dist_in and dist_out."""
    lib = _lib(dist_in.dtype)
    vz = v[2] if len(v) > 2 else None
    _chk(lib.orc_collide_and_propagate(
        ctypes.byref(params), _p(node_map), _p(dist_in), _p(dist_out), _p(rho), _p(v[0]), _p(v[1]),
        _p(vz), ctypes.c_int(options), ctypes.c_uint32(iteration), ctypes.c_int(z_begin),
        ctypes.c_int(z_end)), 'collide_and_propagate')


def set_initial_conditions(params, dist, v, rho, node_map):
    lib = _lib(dist.dtype)
    vz = v[2] if len(v) > 2 else None
    _chk(lib.orc_set_initial_conditions(ctypes.byref(params), _p(dist), _p(v[0]), _p(v[1]), _p(vz),
                                        _p(rho), _p(node_map)), 'set_initial_conditions')


def prepare_macro_fields(params, node_map, dist, rho, options, iteration):
    lib = _lib(dist.dtype)
    _chk(lib.orc_prepare_macro_fields(ctypes.byref(params), _p(node_map), _p(dist), _p(rho),
                                      ctypes.c_int(options), ctypes.c_uint32(iteration)), 'prepare_macro')


def apply_pbc(params, dist, axis, with_swap, links=(0, 0)):
    """links: neighbours across the low / high face of the split axis (see the C function)."""
    lib = _lib(dist.dtype)
    _chk(lib.orc_apply_periodic_boundary_conditions_linked(ctypes.byref(params), _p(dist), int(axis), int(with_swap),
                                                           int(links[0]), int(links[1])), 'apply_pbc')


def collect(params, dist, face, base_gx, base_other, max_lx, max_other, buf, with_swap):
    lib = _lib(dist.dtype)
    _chk(lib.orc_collect_continuous_data(ctypes.byref(params), _p(dist), face, base_gx, base_other,
                                         max_lx, max_other, _p(buf), int(with_swap)), 'collect')


def distribute(params, dist, face, base_gx, base_other, max_lx, max_other, buf, with_swap):
    lib = _lib(dist.dtype)
    _chk(lib.orc_distribute_continuous_data(ctypes.byref(params), _p(dist), face, base_gx, base_other,
                                            max_lx, max_other, _p(buf), int(with_swap)), 'distribute')


def collect_macro(params, field, face, base_gx, base_other, max_lx, max_other, buf):
    lib = _lib(field.dtype)
    _chk(lib.orc_collect_continuous_macro_data(ctypes.byref(params), _p(field), face, base_gx, base_other, max_lx,
                                               max_other, _p(buf)), 'collect_macro')


def distribute_macro(params, field, face, base_gx, base_other, max_lx, max_other, buf):
    lib = _lib(field.dtype)
    _chk(lib.orc_distribute_continuous_macro_data(ctypes.byref(params), _p(field), face, base_gx, base_other, max_lx,
                                                  max_other, _p(buf)), 'distribute_macro')


def lattice_info(lattice_id):
    lib = _lib(np.float32)
    Q, D = ctypes.c_int(), ctypes.c_int()
    e = np.zeros(27 * 3, np.int32)
    w = np.zeros(27, np.float64)
    opp = np.zeros(27, np.int32)
    M = np.zeros(27 * 27, np.int32)
    _chk(lib.orc_lattice_info(lattice_id, ctypes.byref(Q), ctypes.byref(D), _p(e), _p(w), _p(opp), _p(M)),
         'lattice_info')
    q = Q.value
    return dict(Q=q, D=D.value, e=e[:3 * q].reshape(q, 3), w=w[:q], opp=opp[:q], M=M[:q * q].reshape(q, q))


class Simulation(object):
    """Drives the oracle the way SubdomainRunner.step() drives the kernels
    (subdomain_runner.py:960-1058) for a single subdomain: CollideAndPropagate on the
    whole domain, then the per-axis PBC kernels (lb_single.py:153-185,
    subdomain_runner.py:984-1009)."""

    def __init__(self, params, node_map, rho, v, dtype=np.float32):
        self.p = params
        self.dtype = np.dtype(dtype)
        self.map = np.ascontiguousarray(node_map, dtype=np.uint32)
        n = params.dist_size
        self.rho = np.ascontiguousarray(rho, dtype=self.dtype).reshape(-1).copy()
        self.v = [np.ascontiguousarray(c, dtype=self.dtype).reshape(-1).copy() for c in v]
        self.dist_a = np.zeros(params.Q * n, dtype=self.dtype)
        self.dist_b = np.zeros(params.Q * n, dtype=self.dtype) if params.access_pattern == 0 else None
        self.iteration = 0
        set_initial_conditions(params, self.dist_a, self.v, self.rho, self.map)
        if self.dist_b is not None:
            set_initial_conditions(params, self.dist_b, self.v, self.rho, self.map)

    def current_dist(self):
        """Array holding the latest populations (AB ping-pong, lb_single.py:105-106)."""
        if self.dist_b is None or (self.iteration & 1) == 0:
            return self.dist_a
        return self.dist_b

    def step(self, save_macro=False):
        it = self.iteration
        if self.dist_b is None:
            din = dout = self.dist_a
        elif (it & 1) == 0:
            din, dout = self.dist_a, self.dist_b
        else:
            din, dout = self.dist_b, self.dist_a
        collide_and_propagate(self.p, self.map, din, dout, self.rho, self.v,
                              1 if save_macro else 0, it)
        aa = self.dist_b is None
        for axis in range(self.p.dim):
            if self.p.periodic[axis]:
                # _apply_pbc: base = 1 - (iteration & 1); AA even steps use the WithSwap kernel
                apply_pbc(self.p, dout, axis, with_swap=(aa and (it & 1) == 0))
        self.iteration += 1

    def run(self, steps, save_last=True):
        for k in range(steps):
            self.step(save_macro=(save_last and k == steps - 1))


# ---- binary Shan-Chen (BASELINE config 5) ------------------------------------------------------
def sc_set_initial_conditions(params, node_map, d1, d2, v, rho, phi):
    lib = _lib(d1.dtype)
    vz = v[2] if len(v) > 2 else None
    _chk(lib.orc_sc_set_initial_conditions(ctypes.byref(params), _p(node_map), _p(d1), _p(d2), _p(v[0]), _p(v[1]),
                                           _p(vz), _p(rho), _p(phi)), 'sc_set_initial_conditions')


def sc_prepare_macro_fields(params, node_map, d1, d2, rho, phi, v, options, iteration, z_begin=0, z_end=-1):
    lib = _lib(d1.dtype)
    vz = v[2] if len(v) > 2 else None
    _chk(lib.orc_sc_prepare_macro_fields(ctypes.byref(params), _p(node_map), _p(d1), _p(d2), _p(rho), _p(phi),
                                         _p(v[0]), _p(v[1]), _p(vz), int(options), ctypes.c_uint32(iteration),
                                         int(z_begin), int(z_end)), 'sc_prepare_macro_fields')


def sc_collide_and_propagate(params, grid_idx, node_map, din, dout, rho, phi, v, options, iteration,
                             z_begin=0, z_end=-1):
    lib = _lib(din.dtype)
    vz = v[2] if len(v) > 2 else None
    _chk(lib.orc_sc_collide_and_propagate(ctypes.byref(params), int(grid_idx), _p(node_map), _p(din), _p(dout),
                                          _p(rho), _p(phi), _p(v[0]), _p(v[1]), _p(vz), int(options),
                                          ctypes.c_uint32(iteration), int(z_begin), int(z_end)), 'sc_collide')


def apply_macro_pbc(params, field, axis, links=(0, 0)):
    lib = _lib(field.dtype)
    _chk(lib.orc_apply_macro_periodic_boundary_conditions_linked(ctypes.byref(params), _p(field), int(axis),
                                                                 int(links[0]), int(links[1])), 'macro_pbc')


class ShanChenSimulation(object):
    """Step sequence of NNSubdomainRunner.step (subdomain_runner.py:2102-2197) for one subdomain:
    macro fields -> macro PBC -> both collide kernels -> distribution PBC."""

    def __init__(self, params, node_map, rho, phi, v, dtype=np.float32):
        self.p = params
        self.dtype = np.dtype(dtype)
        self.map = np.ascontiguousarray(node_map, dtype=np.uint32)
        n = params.dist_size
        c = lambda a: np.ascontiguousarray(a, dtype=self.dtype).reshape(-1).copy()
        self.rho, self.phi = c(rho), c(phi)
        self.v = [c(x) for x in v]
        ab = params.access_pattern == 0
        self.d = [[np.zeros(params.Q * n, self.dtype), np.zeros(params.Q * n, self.dtype) if ab else None]
                  for _ in range(2)]
        self.iteration = 0
        sc_set_initial_conditions(params, self.map, self.d[0][0], self.d[1][0], self.v, self.rho, self.phi)
        if ab:
            sc_set_initial_conditions(params, self.map, self.d[0][1], self.d[1][1], self.v, self.rho, self.phi)

    def _io(self, g):
        a, b = self.d[g]
        if b is None:
            return a, a
        return (a, b) if (self.iteration & 1) == 0 else (b, a)

    def current(self, g):
        a, b = self.d[g]
        return a if (b is None or (self.iteration & 1) == 0) else b

    def macro(self):
        it = self.iteration
        sc_prepare_macro_fields(self.p, self.map, self._io(0)[0], self._io(1)[0], self.rho, self.phi, self.v, 0, it)
        for axis in range(self.p.dim):
            if self.p.periodic[axis]:
                apply_macro_pbc(self.p, self.rho, axis)
                apply_macro_pbc(self.p, self.phi, axis)

    def collide(self):
        it = self.iteration
        aa = self.d[0][1] is None
        for g in range(2):
            din, dout = self._io(g)
            sc_collide_and_propagate(self.p, g, self.map, din, dout, self.rho, self.phi, self.v, 0, it)
        for axis in range(self.p.dim):
            if self.p.periodic[axis]:
                for g in range(2):
                    apply_pbc(self.p, self._io(g)[1], axis, with_swap=(aa and (it & 1) == 0))

    def step(self):
        self.macro()
        self.collide()
        self.iteration += 1

    def run(self, steps):
        for _ in range(steps):
            self.step()
