"""Worker for tests/test_multi_rank_gloo.py: one rank = one subdomain (the torchrun process
model of sailfish_b200.controller), CPU oracle as the compute, planes exchanged with
torch.distributed (gloo) point-to-point messages."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import multi_util  # noqa: E402
from oracle import oracle  # noqa: E402
from sailfish_b200.geo import EqualSubdomainsGeometry3D  # noqa: E402
from sailfish_b200.examples import ldc_3d, poiseuille_3d  # noqa: E402

CASES = {
    'ldc3d': (ldc_3d.LDCSim, ['--lat_nx=10', '--lat_ny=9', '--lat_nz=14', '--visc=0.04']),
    'poi3d_force_z': (poiseuille_3d.PoiseuilleSim,
                      ['--lat_nx=10', '--lat_ny=10', '--lat_nz=12', '--flow_direction=z', '--drive=force']),
}


def main():
    name, ap, out = sys.argv[1], sys.argv[2], sys.argv[3]
    dist.init_process_group('gloo')
    rank, ws = dist.get_rank(), dist.get_world_size()
    cls, args = CASES[name]
    args = args + ['--access_pattern=' + ap, '--subdomains=%d' % ws]
    _, parts = multi_util.build_case(cls, EqualSubdomainsGeometry3D, args)
    mine = parts[rank]
    sub = multi_util.OracleSubdomain(oracle, mine, np.float64)
    steps = 11
    lat_n = dict((p['spec'].id, p['sd'].arr_shape[0]) for p in parts)
    for k in range(steps):
        sub.collide_and_pbc(save=(k == steps - 1))
        reqs, recvs = [], []
        for face, other in sorted(mine['spec']._face_conn.items()):
            slots, dst, planes = sub.outgoing(face, lat_n[other])
            t = torch.from_numpy(np.ascontiguousarray(planes))
            reqs.append(dist.isend(t, other, tag=face))
            # the neighbour sends through its opposite face
            oface = face ^ 1
            import halo_ref as halo
            rslots, _, rdst = halo.face_transfer(mine['grid'], oface, sub.swap, lat_n[other], sub.shape[0])
            buf = torch.empty_like(t)
            reqs.append(dist.irecv(buf, other, tag=oface))
            recvs.append((rslots, rdst, buf))
        for r in reqs:
            r.wait()
        for rslots, rdst, buf in recvs:
            sub.incoming(rslots, rdst, buf.numpy())
        sub.finish()
    spec = mine['spec']
    piece = np.array([sub.o.rho.reshape(sub.shape)[spec._nonghost_slice]] +
                     [c.reshape(sub.shape)[spec._nonghost_slice] for c in sub.o.v])
    gathered = [None] * ws
    dist.all_gather_object(gathered, (spec.location[2], piece))
    if rank == 0:
        gathered.sort(key=lambda t: t[0])
        np.save(out, np.concatenate([g[1] for g in gathered], axis=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
