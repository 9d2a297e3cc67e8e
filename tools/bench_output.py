"""Cost of writing output: ldc_3d D3Q19 BGK AA with .npz output every N steps, fast path
(slf_pack_field + asynchronous copies + writer thread) vs the reference-style path (--sync_output).
  python tools/bench_output.py [--size 256] [--iters 200] [--every 40]"""
import argparse
import glob
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--iters', type=int, default=400)
    ap.add_argument('--every', type=int, default=40)
    a = ap.parse_args()
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.examples.ldc_3d import LDCSim
    S = a.size
    res = {'config': 'ldc_3d D3Q19 BGK AA %d^3, %d steps, output every %d' % (S, a.iters, a.every)}
    for tag, extra in (('warmup', None), ('none', None), ('fast', []), ('sync', ['--sync_output'])):
        d = tempfile.mkdtemp(prefix='slf_out_')
        args = ['--lat_nx=%d' % S, '--lat_ny=%d' % S, '--lat_nz=%d' % S, '--visc=0.03175', '--access_pattern=AA',
                '--max_iters=%d' % a.iters, '--every=%d' % a.every, '--quiet']
        if extra is not None:
            args += ['--output=' + os.path.join(d, 'o')] + extra
        ctrl = LBSimulationController(LDCSim)
        ctrl.setup(args)
        ctrl.sync()
        t0 = time.time()
        ctrl.run_loop()
        dt = time.time() - t0
        r = ctrl.runners[0]
        files = glob.glob(os.path.join(d, 'o.0.*.npz'))
        res[tag] = {'wall_s': round(dt, 3), 'mlups': round(r.num_fluid_nodes * a.iters / dt * 1e-6, 1),
                    'files': len(files), 'bytes': sum(os.path.getsize(f) for f in files)}
        del ctrl, r
        shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(res))


if __name__ == '__main__':
    main()
