"""Worker for tests/test_multi_rank_gloo.py::test_blocks_*: one rank = one block of a regular block decomposition
(the torchrun process model), the CPU oracle as the compute, the dimension-ordered face exchange of
sailfish_b200/csrc/slf_halo.cuh as MESSAGES (torch.distributed / gloo point to point): every rank only ever touches
its own lattice, so what a window rule forgets to send -- or sends twice -- cannot be papered over by shared memory."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import halo_ref  # noqa: E402
import multi_util  # noqa: E402
from oracle import oracle  # noqa: E402
from sailfish_b200.geo import BlockSubdomainsGeometry2D, BlockSubdomainsGeometry3D  # noqa: E402
from sailfish_b200.examples import ldc_2d, ldc_3d  # noqa: E402

CASES = {
    'ldc3d': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D, ['--lat_nx=11', '--lat_ny=10', '--lat_nz=9', '--visc=0.04']),
    'ldc3d_q27': (ldc_3d.LDCSim, BlockSubdomainsGeometry3D,
                  ['--lat_nx=11', '--lat_ny=10', '--lat_nz=9', '--visc=0.04', '--grid=D3Q27']),
    'ldc2d': (ldc_2d.LDCSim, BlockSubdomainsGeometry2D, ['--lat_nx=15', '--lat_ny=12', '--visc=0.05']),
}


def main():
    name, ap, blocks, out = sys.argv[1:5]
    dist.init_process_group('gloo')
    rank, ws = dist.get_rank(), dist.get_world_size()
    cls, geo, args = CASES[name]
    args = args + ['--access_pattern=' + ap, '--blocks=' + blocks]
    _, parts = multi_util.build_case(cls, geo, args)
    assert len(parts) == ws
    mine = parts[rank]
    spec, grid = mine['spec'], mine['grid']
    D = spec.dim
    sub = multi_util.OracleSubdomain(oracle, mine, np.float64)
    lat_of = dict((p['spec'].id, list(reversed(p['sd'].lat_shape))) for p in parts)
    my_lat = lat_of[spec.id]
    axes = sorted(set(f // 2 for f in spec._face_conn))
    steps = 9
    for k in range(steps):
        sub.collide_and_pbc(save=(k == steps - 1))
        lattice = sub.dist_view(sub.dout)
        for axis in axes:                                   # x, then y, then z: forwarded data must have arrived
            reqs, recvs = [], []
            for side in (0, 1):
                face = 2 * axis + side
                if face not in spec._face_conn:
                    continue
                other = spec._face_conn[face]
                # outgoing through my face
                for n, (slot, src, dst) in enumerate(halo_ref.face_windows(grid, my_lat, face, sub.swap, lat_of[other])):
                    t = torch.from_numpy(np.ascontiguousarray(lattice[(slot,) + tuple(src[::-1])]))
                    reqs.append(dist.isend(t, other, tag=100 * face + n))
                # incoming: the neighbour sends through its opposite face; its windows tell where the data land here
                oface = face ^ 1
                for n, (slot, src, dst) in enumerate(halo_ref.face_windows(grid, lat_of[other], oface, sub.swap, my_lat)):
                    buf = torch.empty(lattice[(slot,) + tuple(dst[::-1])].shape, dtype=torch.float64)
                    reqs.append(dist.irecv(buf, other, tag=100 * oface + n))
                    recvs.append((slot, dst, buf))
            for r in reqs:
                r.wait()
            for slot, dst, buf in recvs:
                lattice[(slot,) + tuple(dst[::-1])] = buf.numpy()
        sub.finish()
    sl = spec._nonghost_slice
    piece = np.array([sub.o.rho.reshape(sub.shape)[sl]] + [c.reshape(sub.shape)[sl] for c in sub.o.v])
    gathered = [None] * ws
    dist.all_gather_object(gathered, (tuple(spec.location), tuple(spec.size), piece))
    if rank == 0:
        gsize = [max(l[a] + s[a] for l, s, _ in gathered) for a in range(D)]
        merged = np.full((piece.shape[0],) + tuple(reversed(gsize)), np.nan)
        for loc, size, f in gathered:
            merged[(slice(None),) + tuple(slice(l, l + s) for l, s in zip(reversed(loc), reversed(size)))] = f
        np.save(out, merged)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
