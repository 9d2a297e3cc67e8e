"""torchrun worker: rank r owns subdomain r on GPU r (the production process model), halo planes
travel GPU-to-GPU through CUDA-IPC mappings ordered by device semaphores.  Rank 0 then repeats
the run as a single subdomain and checks bit-identity of the merged fields."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from sailfish_b200.controller import LBSimulationController  # noqa: E402
from sailfish_b200.geo import EqualSubdomainsGeometry3D  # noqa: E402
from sailfish_b200.examples import ldc_3d, poiseuille_3d, sc_separation_3d  # noqa: E402

CASES = {
    'ldc3d': (ldc_3d.LDCSim, ['--lat_nx=40', '--lat_ny=36', '--lat_nz=64', '--visc=0.04']),
    'poi3d_force_z': (poiseuille_3d.PoiseuilleSim,
                      ['--lat_nx=20', '--lat_ny=20', '--lat_nz=32', '--flow_direction=z', '--drive=force']),
    'poi3d_force_x': (poiseuille_3d.PoiseuilleSim,
                      ['--lat_nx=24', '--lat_ny=20', '--lat_nz=32', '--flow_direction=x', '--drive=force']),
    'sc3d': (sc_separation_3d.SeparationSCSim, ['--lat_nx=20', '--lat_ny=18', '--lat_nz=32', '--seed=5']),
}


def custom_state(ctrl):
    """Decomposition-independent initial state for the mixture (the example's own is random)."""
    for r in ctrl.runners:
        if hasattr(r._sim, 'phi'):
            sim, spec = r._sim, r._spec
            hz, hy, hx = np.meshgrid(np.arange(spec.oz, spec.oz + spec.nz), np.arange(spec.ny), np.arange(spec.nx),
                                     indexing='ij')
            sim.rho[:] = 1.0 + 1e-3 * np.sin(0.7 * hx + 1.3 * hy + 0.4 * hz)
            sim.phi[:] = 1.0 + 1e-3 * np.cos(0.5 * hx - 0.9 * hy + 1.1 * hz)


def collect(ctrl):
    out = []
    for r in ctrl.runners:
        r._fields_to_host()
        sl = r._spec._nonghost_slice
        extra = [np.array(r.field_base(r._sim.phi))[sl]] if hasattr(r._sim, 'phi') else []
        out.append((r._spec.location[2], np.array([np.array(r.field_base(r._sim.rho))[sl]] + extra +
                                                   [np.array(r.field_base(c))[sl] for c in r._sim.v])))
    return out


def run(cls, args, steps, single=False):
    """Returns the results of (a) `steps` steps issued one by one from Python and (b) the same run again
    after LBSimulationController.reinitialize(), issued through run_steps() (CUDA-graph replay)."""
    ctrl = LBSimulationController(cls, EqualSubdomainsGeometry3D)
    ctrl.setup(args + ['--quiet'], local=single)
    custom_state(ctrl)
    ctrl.reinitialize(upload=True)
    for k in range(steps):
        ctrl.step_all(k == steps - 1)
    ctrl.sync()
    first = collect(ctrl)
    # the host fields now hold results: start again from the initial conditions, twice (the second
    # restart finds the semaphores where a completed run left them)
    for _ in range(2):
        for r in ctrl.runners:
            r.reset_host_fields()
        custom_state(ctrl)
        ctrl.reinitialize(upload=True)
        ctrl.run_steps(steps, save_last=True)
        ctrl.sync()
    second = collect(ctrl)
    return first, second


def main():
    name, ap = sys.argv[1], sys.argv[2]
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    # gloo only: the ranks exchange IPC handles and results, never tensors; this also lets several
    # ranks share one GPU (cudaIpc works between processes on the same device)
    ndev = torch.cuda.device_count()
    os.environ['LOCAL_RANK'] = str(local_rank % ndev)
    torch.cuda.set_device(local_rank % ndev)
    dist.init_process_group('gloo')
    rank, ws = dist.get_rank(), dist.get_world_size()
    cls, args = CASES[name]
    args = args + ['--access_pattern=' + ap]
    steps = 21
    mine = run(cls, args + ['--subdomains=%d' % ws], steps)
    gathered = [None] * ws
    dist.all_gather_object(gathered, (mine[0][0], mine[1][0]))
    ok = True
    if rank == 0:
        one = run(cls, args + ['--subdomains=1'], steps, single=True)[0][0][1]
        finite = np.isfinite(one)
        for which, label in ((0, 'stepped'), (1, 're-initialised + graph replay')):
            parts = sorted((g[which] for g in gathered), key=lambda t: t[0])
            merged = np.concatenate([g[1] for g in parts], axis=1)
            same = bool(np.array_equal(merged[finite], one[finite])) and bool((np.isfinite(merged) == finite).all())
            print('RESULT', name, ap, 'ranks', ws, label,
                  'bit-identical' if same else 'MISMATCH max %g' % np.nanmax(np.abs(merged - one)))
            ok = ok and same
        print('RESULT', name, ap, 'all', 'bit-identical' if ok else 'MISMATCH')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
