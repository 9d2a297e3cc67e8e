"""MLUPS of the single-GPU BASELINE.json configs (1-3) through the native runner, device-resident,
CUDA events; steps issued from C (slf_run_steps, CUDA-graph replay) and, for comparison, one
Python call per step.  Prints one JSON line per config."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.examples import ldc_2d, ldc_3d, poiseuille_3d, sc_separation_3d

CONFIGS = [
    ('C1 ldc_2d D2Q9 BGK 256^2', ldc_2d.LDCSim, None,
     ['--lat_nx=256', '--lat_ny=256', '--visc=0.0254', '--access_pattern=AA'], 20000, 72),
    ('C1b ldc_2d D2Q9 BGK 1024^2', ldc_2d.LDCSim, None,
     ['--lat_nx=1024', '--lat_ny=1024', '--visc=0.0254', '--access_pattern=AA'], 4000, 72),
    ('C2 ldc_3d D3Q19 BGK 256^3', ldc_3d.LDCSim, None,
     ['--lat_nx=256', '--lat_ny=256', '--lat_nz=256', '--visc=0.03175', '--access_pattern=AA'], 400, 152),
    ('C2m ldc_3d D3Q19 MRT 256^3', ldc_3d.LDCSim, None,
     ['--lat_nx=256', '--lat_ny=256', '--lat_nz=256', '--visc=0.03175', '--access_pattern=AA', '--model=mrt'], 400, 152),
    ('C2q ldc_3d D3Q27 BGK 256^3', ldc_3d.LDCSim, None,
     ['--lat_nx=256', '--lat_ny=256', '--lat_nz=256', '--visc=0.03175', '--access_pattern=AA', '--grid=D3Q27'], 200, 216),
    ('C2ab ldc_3d D3Q19 BGK 256^3 AB', ldc_3d.LDCSim, None,
     ['--lat_nx=256', '--lat_ny=256', '--lat_nz=256', '--visc=0.03175', '--access_pattern=AB'], 400, 152),
    ('C3 poiseuille_3d D3Q19 MRT 128x128x512 pressure', poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
     ['--lat_nx=128', '--lat_ny=128', '--lat_nz=512', '--flow_direction=z', '--drive=pressure',
      '--model=mrt', '--visc=0.1', '--access_pattern=AA'], 1000, 152),      # (no --stationary: see tests/test_gpu_full_size.py)
    ('C3f poiseuille_3d D3Q19 BGK 128x128x512 force+PBC', poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
     ['--lat_nx=128', '--lat_ny=128', '--lat_nz=512', '--flow_direction=z', '--drive=force',
      '--visc=0.1', '--access_pattern=AA'], 1000, 152),
    ('C5 sc_separation_3d binary Shan-Chen D3Q19 256^3 (1 GPU)', sc_separation_3d.SeparationSCSim, EqualSubdomainsGeometry3D,
     ['--lat_nx=256', '--lat_ny=256', '--lat_nz=256', '--access_pattern=AA', '--seed=1234'], 100, 516),
]


def main():
    peak = 6574.5
    try:
        peak = json.load(open('MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass
    only = [a for a in sys.argv[1:] if not a.startswith('-')]
    for name, cls, geo, args, steps, bpl in CONFIGS:
        if only and name.split()[0] not in only:
            continue
        ctrl = LBSimulationController(cls, geo)
        ctrl.setup(args + ['--quiet'])
        r = ctrl.runners[0]
        dev = r._dev
        nodes = r.num_fluid_nodes
        out = {'config': name, 'fluid_nodes': nodes, 'steps': steps}
        for mode in ('c_graph', 'python'):
            if not hasattr(dev, 'run_python_loop'):
                # binary mixture: graph replay is recorded from the runner's own step()
                fn = (lambda n, save_last=False: r.run_steps(n)) if mode == 'c_graph' else (lambda n, save_last=False: dev.run(n))
            else:
                fn = dev.run if mode == 'c_graph' else dev.run_python_loop
            fn(20, save_last=False)
            ctrl.sync()
            e0 = r.backend.make_event(dev.stream, timing=True)
            fn(steps, save_last=False)
            e1 = r.backend.make_event(dev.stream, timing=True)
            ctrl.sync()
            ms = e1.time_since(e0)
            mlups = nodes * steps / (ms * 1e-3) * 1e-6
            out[mode] = {'mlups': round(mlups, 1), 'us_per_step': round(ms * 1e3 / steps, 2),
                         'alg_GBs': round(mlups * 1e6 * bpl * 1e-9, 1), 'frac_of_measured_hbm': round(mlups * 1e6 * bpl * 1e-9 / peak, 3)}
        print(json.dumps(out))
        sys.stdout.flush()
        del ctrl, r, dev


if __name__ == '__main__':
    main()
