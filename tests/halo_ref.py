"""TEST INFRASTRUCTURE: which planes cross a face of a z-split, host-side statement used by the CPU stand-in device
and the oracle-driven multi-subdomain harness (the device statement is sailfish_b200/csrc/slf_halo.cuh, which the
GPU tests check against the oracle's collect / distribute).

For subdomains split along the outermost axis a face message is, per population that
crosses the face (sym.get_interblock_dists, sym.py:829-845), one whole plane of the SoA
lattice.  Layer and slot selection follow the reference's collect/distribute kernels
(kernel_utils.mako:526-794, subdomain_runner.py:488-510):

  after an AB or AA-odd step  (with_swap=False): ghost layer, natural slots
                                                 -> neighbour's first real layer
  after an AA-even step       (with_swap=True) : first real layer, opposite slots
                                                 -> neighbour's ghost layer
"""
import numpy as np


def face_transfer(grid, face, with_swap, lat_n, peer_lat_n):
    """Returns (slots, src_layer, dst_layer) for the local `face` (2*axis + high)."""
    axis, hi = face // 2, face & 1
    sgn = 1 if hi else -1
    slots = [i for i, e in enumerate(grid.basis) if int(e[axis]) * sgn >= 1]
    if with_swap:
        slots = [int(grid.opposite[i]) for i in slots]
        src = lat_n - 2 if hi else 1
        dst = 0 if hi else peer_lat_n - 1
    else:
        src = lat_n - 1 if hi else 0
        dst = 1 if hi else peer_lat_n - 2
    return slots, src, dst


def with_swap_for(access_pattern_aa, iteration_before_step):
    """The AA in-place (even) step is followed by the WithSwap transfer."""
    return bool(access_pattern_aa and (iteration_before_step & 1) == 0)


def extract_planes(dist, slots, layer):
    """dist: [Q, n_outer, ...] array view.  Returns a [len(slots), ...] copy."""
    return np.stack([dist[s, layer] for s in slots])


def insert_planes(dist, slots, layer, planes):
    for k, s in enumerate(slots):
        dist[s, layer] = planes[k]


def face_windows(grid, lat, face, with_swap, peer_lat, links=0x3f):
    """Dimension-ordered exchange through ANY face (subdomains split along x / y too): for every population that
    crosses `face`, (slot, source index, destination index) as tuples over the axes (x, y[, z]) -- ints along the
    face axis, slices along the others.  Host statement of the window rules of csrc/slf_halo.cuh:
      after streaming steps: earlier axes 1 .. n-2, later axes 1+e .. n-2+e;
      after in-place steps : earlier axes 1-e .. n-2-e, later axes 1 .. n-2;
    `links` (bit 2 * axis + side: that face has a neighbour): along an earlier axis the end position on the side the
    population came from is dropped when nobody delivered anything there (domain boundary)."""
    D = len(lat)
    axis, hi = face // 2, face & 1
    sgn = 1 if hi else -1
    n, pn = lat[axis], peer_lat[axis]
    src_layer = (n - 2 if hi else 1) if with_swap else (n - 1 if hi else 0)
    dst_layer = (0 if hi else pn - 1) if with_swap else (1 if hi else pn - 2)
    out = []
    for i, e in enumerate(grid.basis):
        e = [int(c) for c in e]
        if e[axis] * sgn < 1:
            continue
        slot = int(grid.opposite[i]) if with_swap else i
        src, dst = [None] * D, [None] * D
        for ax in range(D):
            if ax == axis:
                src[ax], dst[ax] = src_layer, dst_layer
                continue
            cnt = lat[ax] - 2
            if ax < axis:
                lo = 1 - e[ax] if with_swap else 1
                came_through = 2 * ax + (0 if e[ax] > 0 else 1)
                if e[ax] != 0 and not (links >> came_through) & 1:
                    # positions lo .. lo+cnt-1 minus the one next to / in the ghost column of that side
                    first, last = lo, lo + cnt - 1
                    if e[ax] > 0:
                        first += 1
                    else:
                        last -= 1
                    lo, cnt = first, last - first + 1
            else:
                lo = 1 if with_swap else 1 + e[ax]
            src[ax] = dst[ax] = slice(lo, lo + max(cnt, 0))
        out.append((slot, tuple(src), tuple(dst)))
    return out
