#!/usr/bin/env python
"""Debug aid for tests/test_gpu_plugin_replay.py: replays a recorded C-ABI call stream on the real library and on
the CPU stand-in device in lockstep and reports the first call after which a device buffer differs.
  python tools/replay_diff.py tests/golden/trace_ldc3d_8blocks_AA.0.pkl.gz"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import abi_trace  # noqa: E402
import fake_device  # noqa: E402
from sailfish_b200 import backend_b200 as bk  # noqa: E402


def main():
    ops = abi_trace.load_fixture(sys.argv[1])
    tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
    real = bk.Lib()
    real.call('slf_init', 0)
    fake = fake_device.FakeLib()
    sizes = {}
    rr, rf = abi_trace.Replayer(real), abi_trace.Replayer(fake)
    reported = 0
    for n, op in enumerate(ops):
        rr.step(op)
        rf.step(op)
        name, rec = op
        if name == 'slf_malloc':
            sizes[rec[1][2]] = rec[0][1]
        if name.startswith('slf_memcpy_d2h') or name in ('slf_malloc', 'slf_event_create', 'slf_event_record',
                                                          'slf_stream_create', 'slf_stream_sync', 'slf_plan_create'):
            continue
        real.call('slf_sync')
        for bid, nbytes in sizes.items():
            a = np.zeros(nbytes, np.uint8)
            real.call('slf_memcpy_d2h', a.ctypes.data_as(ctypes.c_void_p), rr.live['buf'][bid], nbytes)
            b = np.frombuffer(ctypes.string_at(rf.live['buf'][bid], nbytes), dtype=np.uint8)
            if nbytes % 4:
                continue
            x, y = a.view(np.float32), b.view(np.float32)
            fin = np.isfinite(y) & np.isfinite(x)
            bad_mask = not np.array_equal(np.isfinite(x), np.isfinite(y))
            scale = float(np.max(np.abs(y[fin]))) if fin.any() else 0.0
            err = float(np.max(np.abs(x[fin] - y[fin])) / scale) if fin.any() and scale > 0 else 0.0
            # integer buffers (maps, index lists) compare exactly
            if (err > tol or bad_mask) and not np.array_equal(a, b):
                k = int(np.argmax(np.abs(np.where(fin, x - y, 0))))
                print('op %d %s %s: buffer %d (%d B) differs: rel %.3g, finite-mask %s, worst at float %d: gpu %r cpu %r'
                      % (n, name, [r if r and r[0] in ('i', 'buf') else (r[0] if r else None) for r in rec], bid, nbytes,
                         err, 'differs' if bad_mask else 'same', k, x[k], y[k]))
                reported += 1
        if reported >= 6:
            break
    print('done: %d differing buffers reported' % reported)


if __name__ == '__main__':
    main()
