"""One rank per GPU (torchrun), CUDA-IPC peer pushes + device semaphores: the merged result must
be bit-identical to the single-subdomain run.  With a single GPU two ranks share it."""
import os
import subprocess
import sys

import pytest

import util

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize('name,ap', [('ldc3d', 'AA'), ('ldc3d', 'AB'), ('poi3d_force_z', 'AA'), ('poi3d_force_x', 'AA'), ('sc3d', 'AA'), ('sc3d', 'AB')])
def test_ranks_equal_single_subdomain(name, ap):
    n = _ngpu()
    if n < 1:
        pytest.skip('needs a GPU')
    # one rank per GPU; on a single-GPU box the ranks share the device (CUDA IPC between processes
    # works on one device too), which still exercises the cross-process path
    n = min(n, 4) if n >= 2 else 2
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % n,
           '--master-addr', '127.0.0.1', '--master-port', str(port),
           os.path.join(util.ROOT, 'tests', '_gpu_rank_worker.py'), name, ap]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-3000:])
    assert 'bit-identical' in res.stdout
