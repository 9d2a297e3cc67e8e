"""BASELINE config 5: binary Shan-Chen D3Q19 256^3 (examples/binary_fluid/sc_separation_3d.py at
--lat_n{x,y,z}=256) split along z over the launched ranks (torchrun), device-resident MLUPS.
  python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 tools/bench_sc.py [--size 256]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--steps', type=int, default=200)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    lr = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(lr)
    if ws > 1:
        dist.init_process_group('cpu:gloo,cuda:nccl', device_id=torch.device('cuda', lr))
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.geo import EqualSubdomainsGeometry3D
    from sailfish_b200.examples.sc_separation_3d import SeparationSCSim
    ctrl = LBSimulationController(SeparationSCSim, EqualSubdomainsGeometry3D)
    S = a.size
    ctrl.setup(['--lat_nx=%d' % S, '--lat_ny=%d' % S, '--lat_nz=%d' % S, '--subdomains=%d' % ws,
                '--access_pattern=AA', '--seed=1234', '--quiet'])
    r = ctrl.runners[0]

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctrl.sync()
    r.run_steps(10)                 # records the step-pair graph
    barrier()
    e0 = r.backend.make_event(r._calc_stream, timing=True)
    r.run_steps(a.steps)            # CUDA-graph replay of step pairs recorded from NNSubdomainRunner.step
    e1 = r.backend.make_event(r._calc_stream, timing=True)
    barrier()
    ms = e1.time_since(e0)
    if ws > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    r._fields_to_host()
    import numpy as np
    ok = bool(np.isfinite(np.asarray(r.field_base(r._sim.rho))[r._spec._nonghost_slice]).all())
    if int(os.environ.get('RANK', '0')) == 0:
        nodes = S ** 3
        mlups = nodes * a.steps / (ms * 1e-3) * 1e-6
        print(json.dumps({'config': 'C5 binary Shan-Chen D3Q19 %d^3, %d GPU(s), z-split, AA' % (S, ws),
                          'mlups': mlups, 'ms_per_step': ms / a.steps,
                          'alg_GBs_at_516B_per_LUP': mlups * 1e6 * 516e-9 / max(ws, 1),
                          'finite': ok}))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
