"""torchrun worker: 8 ranks = the 2 x 2 x 2 blocks of a lid-driven cavity (regtest/subdomains/3d_ldc.py layout),
ranks spread round-robin over the visible GPUs; faces exchanged GPU to GPU (CUDA IPC + device semaphores).  Rank 0
repeats the run as one subdomain and checks bit-identity of the merged fields -- stepped one by one, and again after a
re-initialisation through run_steps() (CUDA-graph replay)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import block_util  # noqa: E402
from sailfish_b200.controller import LBSimulationController  # noqa: E402
from sailfish_b200.examples import ldc_3d  # noqa: E402
from sailfish_b200.geo import BlockSubdomainsGeometry3D  # noqa: E402

GSIZE = (40, 36, 34)
ARGS = ['--lat_nx=40', '--lat_ny=36', '--lat_nz=34', '--visc=0.04']


def main():
    ap = sys.argv[1]
    blocks = sys.argv[2] if len(sys.argv) > 2 else '2,2,2'
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    ndev = torch.cuda.device_count()
    os.environ['LOCAL_RANK'] = str(local_rank % ndev)
    torch.cuda.set_device(local_rank % ndev)
    dist.init_process_group('gloo')
    rank, ws = dist.get_rank(), dist.get_world_size()
    args = ARGS + ['--access_pattern=' + ap, '--quiet']
    steps = 21
    ctrl = LBSimulationController(ldc_3d.LDCSim, BlockSubdomainsGeometry3D)
    ctrl.setup(args + ['--blocks=' + blocks])
    for k in range(steps):
        ctrl.step_all(k == steps - 1)
    ctrl.sync()
    first = block_util.collect(ctrl)[0]
    ctrl.reinitialize(reset_fields=True)
    ctrl.run_steps(steps, save_last=True)
    ctrl.sync()
    second = block_util.collect(ctrl)[0]
    gathered = [None] * ws
    dist.all_gather_object(gathered, (first, second))
    ok = True
    if rank == 0:
        _, one = block_util.run_blocks(ldc_3d.LDCSim, BlockSubdomainsGeometry3D, args + ['--blocks=1,1,1'], steps, local=True)
        ref = block_util.merge(one, GSIZE)
        for which, label in ((0, 'stepped'), (1, 're-initialised + graph replay')):
            got = block_util.merge([g[which] for g in gathered], GSIZE)
            same = bool(np.array_equal(got, ref, equal_nan=True))
            print('RESULT', ap, 'blocks', blocks, 'ranks', ws, 'gpus', ndev, label, 'bit-identical' if same else
                  'MISMATCH max %g' % np.nanmax(np.abs(got - ref)))
            ok = ok and same
        print('RESULT', ap, blocks, 'all bit-identical' if ok else 'MISMATCH')
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
