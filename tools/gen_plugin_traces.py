#!/usr/bin/env python
"""Records the C-ABI call stream of the UNMODIFIED reference driving the b200 backend (plugin mode) on the CPU
stand-in device and writes it as fixtures tests/golden/trace_*.pkl.gz (see tests/abi_trace.py).  Build
container only: needs /root/reference.  `python tools/gen_plugin_traces.py` regenerates all fixtures;
`python tools/gen_plugin_traces.py --child <case> <dir>` is the recording process itself."""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('SAILFISH_REFERENCE', '/root/reference')
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tools')):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {
    # name: (example script, class, geometry, args, steps, every)
    'ldc3d_AA': ('examples/ldc_3d.py', 'LDCSim', None,
                 ['--lat_nx=12', '--lat_ny=10', '--lat_nz=11', '--visc=0.04', '--access_pattern=AA'], 10, 5),
    'ldc2d_mrt_AB': ('examples/ldc_2d.py', 'LDCSim', None,
                     ['--lat_nx=16', '--lat_ny=14', '--visc=0.05', '--model=mrt', '--access_pattern=AB'], 10, 5),
    'poi3d_force_AB': ('examples/poiseuille_3d.py', 'PoiseuilleSim', 'EqualSubdomainsGeometry3D',
                       ['--grid=D3Q19', '--lat_nx=14', '--lat_ny=12', '--lat_nz=12', '--flow_direction=x', '--drive=force',
                        '--visc=0.1', '--access_pattern=AB'], 10, 5),
    'poi3d_press_mrt_AA': ('examples/poiseuille_3d.py', 'PoiseuilleSim', 'EqualSubdomainsGeometry3D',
                           ['--grid=D3Q19', '--lat_nx=12', '--lat_ny=12', '--lat_nz=16', '--flow_direction=z',
                            '--drive=pressure', '--stationary', '--model=mrt', '--visc=0.1', '--access_pattern=AA'], 10, 5),
    'sc3d_AA': ('examples/binary_fluid/sc_separation_3d.py', 'SeparationSCSim', 'LBGeometry3D',
                ['--lat_nx=12', '--lat_ny=10', '--lat_nz=9', '--seed=1234', '--access_pattern=AA'], 8, 4),
    # 2 x 2 x 2 blocks: continuous y/z-face transfers + sparse x-face / edge / corner transfers, one process each
    'ldc3d_8blocks_AA': ('examples/ldc_3d.py', 'LDCSim', '2x2x2',
                         ['--lat_nx=16', '--lat_ny=14', '--lat_nz=12', '--visc=0.04', '--access_pattern=AA'], 8, 4),
    'ldc3d_8blocks_AB': ('examples/ldc_3d.py', 'LDCSim', '2x2x2',
                         ['--lat_nx=16', '--lat_ny=14', '--lat_nz=12', '--visc=0.04', '--access_pattern=AB'], 6, 3),
    # tests/geometry_zoo.py with the reference's classes: node classes the examples do not use
    'zoo_tms_block3d_AA': ('zoo:tms_block3d', None, None, ['--visc=0.06', '--access_pattern=AA'], 8, 4),
    'zoo_copy_duct3d_AB': ('zoo:copy_duct3d', None, None, ['--visc=0.06', '--access_pattern=AB'], 8, 4),
    'zoo_slip_box3d_mrt_AA': ('zoo:slip_box3d_mrt', None, None, ['--visc=0.06', '--access_pattern=AA'], 8, 4),
}


def child(case, out_dir):
    from sailfish_b200 import plugin
    plugin.install(REF, default=True)
    sys.path.insert(0, REF)
    import fake_device
    import abi_trace
    from oracle import oracle as orc
    orc.build()
    fake = fake_device.install_fake()
    abi_trace.Recorder(fake, out_dir)
    import ref_harness
    script, cls_name, geo_name, args, steps, every = CASES[case]
    from sailfish.controller import LBSimulationController
    from sailfish import geo
    if script.startswith('zoo:'):
        import geometry_zoo
        from sailfish import node_type as rnt
        from sailfish.subdomain import Subdomain2D, Subdomain3D
        from sailfish.lb_single import LBFluidSim
        from sailfish.lb_base import LBForcedSim
        ns = dict(Subdomain2D=Subdomain2D, Subdomain3D=Subdomain3D, LBFluidSim=LBFluidSim, LBForcedSim=LBForcedSim)
        ns.update((n, getattr(rnt, n)) for n in dir(rnt) if n.startswith('NT'))
        sim_cls, zoo_args = geometry_zoo.make_cases(ns)[script[len('zoo:'):]]
        d = {}
        sim_cls.update_defaults(d)
        out = os.path.join(out_dir, 'out')
        sys.argv = [script] + zoo_args + args + ['--debug_single_process', '--backends=b200', '--max_iters=%d' % steps,
                                                 '--every=%d' % every, '--output=' + out, '--quiet']
        LBSimulationController(sim_cls, geo.LBGeometry2D if d['grid'] == 'D2Q9' else geo.LBGeometry3D).run()
        return
    mod = ref_harness.load_example(script)
    if geo_name == '2x2x2':
        from sailfish.subdomain import SubdomainSpec3D

        class Blocks(geo.LBGeometry3D):
            def subdomains(self, n=None):
                out = []
                xq, yq, zq = self.gx // 2, self.gy // 2, self.gz // 2
                for i in range(2):
                    for j in range(2):
                        for k in range(2):
                            out.append(SubdomainSpec3D((i * xq, j * yq, k * zq),
                                                       (self.gx - xq if i else xq, self.gy - yq if j else yq,
                                                        self.gz - zq if k else zq)))
                return out
        geo_cls = Blocks
        extra = []
    else:
        geo_cls = getattr(geo, geo_name) if geo_name else None
        extra = ['--debug_single_process']
    out = os.path.join(out_dir, 'out')
    sys.argv = [script] + args + extra + ['--backends=b200', '--max_iters=%d' % steps, '--every=%d' % every,
                                          '--output=' + out, '--quiet']
    LBSimulationController(getattr(mod, cls_name), geo_cls).run()


def main():
    import abi_trace
    gold = os.path.join(ROOT, 'tests', 'golden')
    only = sys.argv[1:]
    for case in CASES:
        if only and case not in only:
            continue
        d = tempfile.mkdtemp()
        try:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), '--child', case, d],
                                  env=dict(os.environ, OMP_NUM_THREADS='1'))
            files = sorted(glob.glob(os.path.join(d, 'trace_*.pkl')))
            traces = [abi_trace.load(f) for f in files]
            traces = [t for t in traces if any(op[0] == 'slf_plan_create' for op in t)]
            for old in glob.glob(os.path.join(gold, 'trace_%s.*.pkl.gz' % case)):
                os.unlink(old)
            for k, t in enumerate(traces):
                path = os.path.join(gold, 'trace_%s.%d.pkl.gz' % (case, k))
                abi_trace.save_fixture(t, path)
                nd2h = sum(1 for op in t for a in op[1] if a and a[0] == 'd2h')
                print(case, k, len(t), 'ops', nd2h, 'device->host copies', os.path.getsize(path), 'bytes')
        finally:
            shutil.rmtree(d, ignore_errors=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--child':
        child(sys.argv[2], sys.argv[3])
    else:
        main()
