"""Block decompositions (subdomains split along x and / or y too) on the GPU: dimension-ordered face exchange
(slf_halo_push_face) with edge / corner data travelling in two or three hops.
  * one process driving all blocks on one GPU (CUDA events between the runners' streams);
  * one rank per block under torchrun (CUDA-IPC peer stores + device semaphores, graph replay), ranks spread
    round-robin over the visible GPUs -- regtest/subdomains/3d_ldc.py's 8-block layout.
The merged fields must equal the single-subdomain run bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

import block_util
import util
from sailfish_b200.examples import ldc_2d, ldc_3d
from sailfish_b200.geo import BlockSubdomainsGeometry2D, BlockSubdomainsGeometry3D, EqualSubdomainsGeometry3D

pytestmark = pytest.mark.gpu

L3 = ['--lat_nx=40', '--lat_ny=36', '--lat_nz=34', '--visc=0.04']
CASES = {
    'ldc3d_2x2x2': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3, ['--blocks=2,2,2'], (40, 36, 34)),
    'ldc3d_q27_2x2x2': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3 + ['--grid=D3Q27'], ['--blocks=2,2,2'], (40, 36, 34)),
    'ldc3d_mrt_3x2x1': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, L3 + ['--model=mrt'], ['--blocks=3,2,1'], (40, 36, 34)),
    'ldc3d_conn_x': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D, L3, ['--subdomains=3', '--conn_axis=x'], (40, 36, 34)),
    'ldc2d_2x2': (ldc_2d.LDCSim, BlockSubdomainsGeometry2D, ['--lat_nx=70', '--lat_ny=50', '--visc=0.05'], ['--blocks=2,2'], (70, 50)),
}


@pytest.mark.parametrize('ap', ['AA', 'AB'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_blocks_equal_single_one_process(name, ap):
    cls, geo, args, split, gsize = CASES[name]
    args = args + ['--access_pattern=' + ap]
    single = ['--blocks=' + ','.join(['1'] * len(gsize))] if any(a.startswith('--blocks') for a in split) else ['--subdomains=1']
    _, one = block_util.run_blocks(cls, geo, args + single, 21)
    _, many = block_util.run_blocks(cls, geo, args + split, 21)
    np.testing.assert_array_equal(block_util.merge(one, gsize), block_util.merge(many, gsize))


@pytest.mark.parametrize('ap', ['AA', 'AB'])
def test_eight_ranks_equal_single(ap):
    import torch
    if torch.cuda.device_count() < 1:
        pytest.skip('needs a GPU')
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=8', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(util.ROOT, 'tests', '_gpu_block_worker.py'), ap]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-3000:])
    assert 'all bit-identical' in res.stdout
