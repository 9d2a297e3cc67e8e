"""Output fast path.

The reference brings rho and v to the host with synchronous copies of the padded, ghost-inclusive
arrays (``SubdomainRunner._fields_to_host``, sailfish/subdomain_runner.py:1141-1151), masks the
non-wet nodes on the host (``LBOutput.mask_nonfluid_nodes``, sailfish/io.py:53-59) and hands the
arrays to a writer thread (``NPYOutput``, sailfish/io.py:294-348).  With a step of a few hundred
microseconds all of that is on the critical path.  Here:

* ``slf_pack_field`` strips ghosts/padding and writes the NaN mask on the device (one pass at HBM
  speed, on the compute stream: it is ordered after the step that produced the fields and before the
  next step that overwrites them);
* the dense arrays go to pinned host memory with asynchronous copies on a separate stream, triple
  buffered, while the simulation continues;
* writer threads wait for the copy's event and calls ``numpy.savez`` -- same file name scheme and
  same array names / shapes as ``LBOutput.save``.
"""
import queue
import threading

import numpy as np

from sailfish_b200 import backend_b200 as bk


class FastOutput(object):
    def __init__(self, runner, output, slots=3, writers=2):
        self.runner = runner
        self.output = output
        self.be = runner.backend
        self.lib = bk.lib()
        self.device = getattr(self.be, 'gpu_id', 0)
        self.fields = []          # (name, is_vector, [device buffers])
        for name, view in output._scalar.items():
            self.fields.append((name, False, [runner.gpu_field(view)]))
        for name, comps in output._vector.items():
            self.fields.append((name, True, list(runner.gpu_field(comps))))
        spec = runner._spec
        self.shape = tuple(runner._lat_size[i] - 2 for i in range(len(runner._lat_size)))
        n = int(np.prod(self.shape))
        self.nbytes = n * 4
        ncomp = sum(len(b) for _, _, b in self.fields)
        self.slots = []
        for _ in range(slots):
            self.slots.append({
                'dev': [self.be.alloc_buf(size=self.nbytes) for _ in range(ncomp)],
                'host': [self.be.alloc_async_host_buf(self.shape, np.float32) for _ in range(ncomp)],
                'free': threading.Event(),
            })
            self.slots[-1]['free'].set()
        self.copy_stream = self.be.make_stream()
        self.packed = self.be.make_event(runner._calc_stream)
        self.k = 0
        self.stall_s = 0.0        # time the main loop spent waiting for a free slot
        self.errors = []
        self.q = queue.Queue()
        # savez spends its time in crc32 + write(), both of which release the GIL: two writers overlap files
        self.threads = [threading.Thread(target=self._writer, name='slf-output-%d' % i, daemon=True)
                        for i in range(writers)]
        for t in self.threads:
            t.start()

    # -- main thread -----------------------------------------------------------------------------
    def save(self, it):
        """Called right after the step that stored the fields was issued on the compute stream."""
        import time
        r = self.runner
        slot = self.slots[self.k % len(self.slots)]
        self.k += 1
        t0 = time.time()
        slot['free'].wait()
        self.stall_s += time.time() - t0
        slot['free'].clear()
        cs = r._calc_stream
        d_map = int(r.gpu_geo_map())
        j = 0
        for _, _, bufs in self.fields:
            for b in bufs:
                self.lib.call('slf_pack_field', r._dev.plan, d_map, int(b), int(slot['dev'][j]), cs.handle)
                j += 1
        self.packed.record(cs)
        self.copy_stream.wait_for_event(self.packed)
        for j, (d, h) in enumerate(zip(slot['dev'], slot['host'])):
            self.lib.call('slf_memcpy_d2h_async', h.ctypes.data_as(bk._vp), int(d), int(self.nbytes),
                          self.copy_stream.handle)
        done = self.be.make_event(self.copy_stream)
        self.q.put((it, slot, done))

    def close(self):
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.errors:
            raise self.errors[0]

    # -- writer thread ---------------------------------------------------------------------------
    def _writer(self):
        try:
            self.lib.call('slf_init', int(self.device))
        except Exception as e:      # pragma: no cover
            self.errors.append(e)
        while True:
            item = self.q.get()
            if item is None:
                return
            it, slot, done = item
            try:
                done.synchronize()
                if self.output.basename:
                    data, j = {}, 0
                    for name, is_vec, bufs in self.fields:
                        if is_vec:
                            data[name] = np.array(slot['host'][j:j + len(bufs)])
                        else:
                            data[name] = slot['host'][j]
                        j += len(bufs)
                    np.savez(self.output.filename(it), **data)
            except Exception as e:
                self.errors.append(e)
            finally:
                slot['free'].set()
