"""Record / replay of the C-ABI call stream (TEST INFRASTRUCTURE).

The unmodified reference can only drive the backend where the reference tree exists (the build container,
no GPU), and the GPU box has no reference tree.  So the calls the reference's own runner makes through
``sailfish_b200.backend_b200`` are RECORDED here on the CPU stand-in device (tests/fake_device.py: host
memory + the oracle) -- every entry point with its arguments, every host -> device payload, and what every
device -> host copy returned -- and REPLAYED on the GPU box against the real ``libslf_b200.so``: same calls,
same order, same payloads; every device -> host copy is compared with what the oracle-backed device
returned (tests/test_gpu_plugin_replay.py).  A trace is self-contained per process: ghost-layer data
arriving from another subdomain's process is just another recorded host -> device payload.

Trace format: a list of ops ``(name, [arg, ...])`` with args
    ('i', int) | ('null',) | ('buf', id, offset) | ('plan', id) | ('stream', id) | ('event', id) | ('graph', id)
    ('h2d', bytes) | ('d2h', nbytes, bytes_expected) | ('params', struct_bytes, [node_params]) | ('new', kind) | ('skip',)
"""
import ctypes
import os
import pickle

import numpy as np

from sailfish_b200 import backend_b200 as bk
from sailfish_b200 import params as P

_NEW_KIND = {'slf_malloc': 'buf', 'slf_stream_create': 'stream', 'slf_event_create': 'event',
             'slf_plan_create': 'plan', 'slf_host_alloc': 'host', 'slf_graph_end_capture': 'graph'}
# argument kinds of the entry points that are not kernels; kernels follow the default rule of `signature`
_SIG = {
    'slf_malloc': ['size', 'new'], 'slf_free': ['buf'], 'slf_memset': ['buf', 'i', 'size', 'stream'],
    'slf_host_alloc': ['size', 'new'], 'slf_memcpy_h2d': ['buf', 'h2d', 'size'], 'slf_memcpy_d2h': ['d2h', 'buf', 'size'],
    'slf_memcpy_h2d_async': ['buf', 'h2d', 'size', 'stream'], 'slf_memcpy_d2h_async': ['d2h', 'buf', 'size', 'stream'],
    'slf_memcpy_d2d_async': ['buf', 'buf', 'size', 'stream'], 'slf_sync': [],
    'slf_stream_create': ['new', 'i'], 'slf_stream_destroy': ['stream'], 'slf_stream_sync': ['stream'],
    'slf_stream_wait_event': ['stream', 'event'], 'slf_event_create': ['new', 'i'], 'slf_event_destroy': ['event'],
    'slf_event_record': ['event', 'stream'], 'slf_event_sync': ['event'], 'slf_event_elapsed_ms': ['event', 'event', 'skip'],
    'slf_plan_create': ['params', 'new'], 'slf_plan_destroy': ['plan'], 'slf_plan_set_links': ['plan', 'i', 'i'],
    'slf_plan_bind_map': ['plan', 'buf'], 'slf_graph_begin_capture': ['stream'], 'slf_graph_end_capture': ['stream', 'new'],
    'slf_graph_launch': ['graph', 'stream'], 'slf_graph_destroy': ['graph'],
}
_SKIP = {'slf_device_info', 'slf_device_count', 'slf_init', 'slf_host_free', 'slf_ipc_get_handle',
         'slf_ipc_open_handle', 'slf_ipc_close_handle', 'slf_enable_peer_access', 'slf_sem_status', 'slf_sem_abort'}


def signature(name):
    """Kinds of the arguments of entry point `name`.  Kernels: the plan, then device pointers (uint64) and
    integers, the stream last."""
    if name in _SIG:
        return _SIG[name]
    types = bk._PROTOTYPES[name]
    sig = []
    for k, t in enumerate(types):
        if t is bk._vp:
            sig.append('plan')
        elif t is bk._u64:
            sig.append('stream' if k == len(types) - 1 else 'buf')
        else:
            sig.append('i')
    return sig


def _addr(a):
    if a is None:
        return 0
    if isinstance(a, int):
        return int(a)
    if hasattr(a, 'value'):
        return int(a.value or 0)
    return int(ctypes.cast(a, ctypes.c_void_p).value or 0)


class Recorder(object):
    """Wraps FakeLib.call.  Ops are appended to ``<dir>/trace_<pid>.pkl`` as they happen (the reference runs
    every subdomain in a forked process that exits through os._exit)."""

    def __init__(self, fake, out_dir):
        self.fake = fake
        self.dir = out_dir
        self._pid = None
        self._f = None
        self._ids = {}            # kind -> {handle value: id}
        self._stock = fake.call
        fake.call = self.call

    def _file(self):
        if self._pid != os.getpid():
            self._pid = os.getpid()
            self._f = open(os.path.join(self.dir, 'trace_%d.pkl' % self._pid), 'ab')
        return self._f

    def _emit(self, op):
        f = self._file()
        pickle.dump(op, f, protocol=4)
        f.flush()

    def _id(self, kind, value, create=False):
        tab = self._ids.setdefault(kind, {})
        if create or value not in tab:
            tab[value] = len(tab) + 1 if not create else max(list(tab.values()) + [0]) + 1
        return tab[value]

    def _buf(self, value):
        """(id, offset) of the device buffer that contains address `value`."""
        for base, arr in self.fake.bufs.items():
            if base <= value < base + arr.nbytes:
                return self._id('buf', base), value - base
        raise KeyError('device address 0x%x is not inside a recorded allocation' % value)

    def call(self, name, *args):
        if name in _SKIP or self.fake._capture is not None or name not in bk._PROTOTYPES:
            return self._stock(name, *args)
        sig = signature(name)
        rec, post = [], []
        for k, (a, kind) in enumerate(zip(args, sig)):
            if kind in ('i', 'size'):
                rec.append(('i', int(a)))
            elif kind == 'buf':
                v = _addr(a)
                rec.append(('null',) if v == 0 else ('buf',) + self._buf(v))
            elif kind in ('stream', 'event', 'graph', 'plan'):
                v = _addr(a)
                rec.append(('null',) if v == 0 and kind == 'stream' else (kind, self._id(kind, v)))
            elif kind == 'h2d':
                rec.append(('h2d', ctypes.string_at(_addr(a), int(args[2]))))
            elif kind == 'd2h':
                rec.append(None)
                post.append((len(rec) - 1, _addr(a), int(args[2])))
            elif kind == 'params':
                p = a._obj
                n = int(p.num_node_params)
                np_list = [float(p.node_params[j]) for j in range(n)] if n > 0 else []
                rec.append(('params', ctypes.string_at(ctypes.addressof(p), ctypes.sizeof(p)), np_list))
            elif kind == 'new':
                rec.append(None)
                post.append((len(rec) - 1, a, 'new'))
            else:
                rec.append(('skip',))
        rc = self._stock(name, *args)
        for pos, obj, what in post:
            if what == 'new':
                kind = _NEW_KIND[name]
                val = int(obj._obj.value or 0)
                # the stand-in device hands out the same stream / event handle every time: one id each
                rec[pos] = ('new', kind, self._id(kind, val, create=kind not in ('stream', 'event')))
            else:
                rec[pos] = ('d2h', what, ctypes.string_at(obj, what))
        self._emit((name, rec))
        return rc


def load(path):
    ops = []
    with open(path, 'rb') as f:
        while True:
            try:
                ops.append(pickle.load(f))
            except EOFError:
                break
    return ops


def save_fixture(ops, path):
    with open(path, 'wb') as f:
        import gzip
        with gzip.GzipFile(fileobj=f, mode='wb', mtime=0) as g:
            pickle.dump(ops, g, protocol=4)


def load_fixture(path):
    import gzip
    with gzip.open(path, 'rb') as g:
        return pickle.load(g)


class Replayer(object):
    """Issues recorded calls on `lib` (a backend_b200.Lib or FakeLib), one op at a time."""

    def __init__(self, lib):
        self.lib = lib
        self.live = {'buf': {}, 'plan': {}, 'stream': {}, 'event': {}, 'graph': {}, 'host': {}}
        self.keep = []                  # host payloads must outlive asynchronous copies
        self.results = []
        self.n = 0

    def handle(self, kind, i):
        live, lib = self.live, self.lib
        if kind in ('stream', 'event') and i not in live[kind]:
            h = ctypes.c_uint64()
            if kind == 'stream':
                lib.call('slf_stream_create', ctypes.byref(h), 0)
            else:
                lib.call('slf_event_create', ctypes.byref(h), 0)
            live[kind][i] = h.value
        return live[kind][i]

    def step(self, op):
        name, rec = op
        lib, live, keep = self.lib, self.live, self.keep
        n = self.n
        self.n += 1
        args, outs, checks = [], [], []
        for r in rec:
            k = r[0]
            if k == 'i':
                args.append(r[1])
            elif k == 'null':
                args.append(0)
            elif k == 'buf':
                args.append(live['buf'][r[1]] + r[2])
            elif k == 'plan':
                args.append(ctypes.c_void_p(live['plan'][r[1]]))
            elif k in ('stream', 'event'):
                args.append(self.handle(k, r[1]))
            elif k == 'graph':
                args.append(live['graph'][r[1]])
            elif k == 'h2d':
                a = np.frombuffer(r[1], dtype=np.uint8).copy()
                keep.append(a)
                args.append(a.ctypes.data_as(ctypes.c_void_p))
            elif k == 'd2h':
                a = np.zeros(r[1], dtype=np.uint8)
                keep.append(a)
                args.append(a.ctypes.data_as(ctypes.c_void_p))
                checks.append((a, r[2]))
            elif k == 'params':
                p = P.SlfParams.from_buffer_copy(r[1])
                arr = (ctypes.c_double * max(1, len(r[2])))(*r[2])
                p.node_params = ctypes.cast(arr, ctypes.POINTER(ctypes.c_double))
                keep.append((p, arr))
                args.append(ctypes.byref(p))
            elif k == 'new':
                kind = r[1]
                if kind in ('stream', 'event'):
                    args.append(None)
                    outs.append(None)
                    continue
                h = ctypes.c_void_p() if kind in ('plan', 'host') else ctypes.c_uint64()
                args.append(ctypes.byref(h))
                outs.append((kind, r[2], h))
            else:
                args.append(None)
        if name in ('slf_stream_create', 'slf_event_create'):
            # created lazily by handle(): the stand-in device hands out one handle value for all of them
            self.handle('stream' if name == 'slf_stream_create' else 'event', rec[0][2])
            return
        if name in ('slf_stream_destroy', 'slf_event_destroy', 'slf_event_elapsed_ms', 'slf_host_alloc', 'slf_free',
                    'slf_plan_destroy'):
            return                  # lifetime management of the recording process; everything lives until the end here
        lib.call(name, *args)
        for o in outs:
            if o is not None:
                live[o[0]][o[1]] = int(o[2].value or 0)
        if checks:
            lib.call('slf_sync')
            for got, want in checks:
                g = np.frombuffer(got.tobytes(), dtype=np.float32)
                w = np.frombuffer(want, dtype=np.float32)
                fin = np.isfinite(w)
                mask_ok = bool(np.array_equal(np.isfinite(g), fin))
                # relative to the largest value of the copy, but never below 1e-2: a velocity component that is
                # physically zero in this block holds round-off noise of populations that are O(0.01 .. 1)
                scale = max(float(np.max(np.abs(w[fin]))), 1e-2) if fin.any() else 1.0
                err = float(np.max(np.abs(g[fin] - w[fin])) / scale) if fin.any() and mask_ok else 0.0
                self.results.append((n, name, err, mask_ok, int(fin.sum())))


def replay(ops, lib):
    """Issues the recorded calls on `lib`.  Returns the list of comparisons (op index, name, max relative
    error over finite values, whether the finite masks agree, number of finite values)."""
    r = Replayer(lib)
    for op in ops:
        r.step(op)
    lib.call('slf_sync')
    return r.results
