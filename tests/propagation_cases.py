"""Literal known-answer vectors of the reference's streaming regtest
regtest/subdomains/3d_propagation.py (TwoBlockPropagationTest, `axis = 2` / DepthTest variant,
:120-176): two 64 x 66 x 64 blocks connected along z, collision disabled
(`config.relaxation_enabled = False`, :60-62), single populations placed by hand, and the exact
(slot, z, y, x) where each value must be after ONE step (AB pattern).  Coordinates are
ghost-inclusive array indices, as in the reference file.

The reference writes them through the permutations p1/p2 (:99-117); for axis == 2:
    p1(x, y, z) = (y, z, x)           direction vector
    p2(d, z, y, x) = (d, x, z, y)     index into dbuf[dist, z, y, x]
"""
from sailfish_b200 import lattice
from sailfish_b200.geo import EqualSubdomainsGeometry3D
from sailfish_b200.lb_single import LBFluidSim
from sailfish_b200.subdomain import Subdomain3D

L = lattice.get('D3Q19')


def vi(x, y, z):
    return int([i for i, e in enumerate(L.basis) if tuple(e) == (x, y, z)][0])


def p1(x, y, z):
    return (y, z, x)


def p2(d, z, y, x):
    return (d, x, z, y)


class BlockTest(Subdomain3D):
    def boundary_conditions(self, hx, hy, hz):
        pass

    def initial_conditions(self, sim, hx, hy, hz):
        pass


class SimulationTest(LBFluidSim):
    subdomain = BlockTest

    @classmethod
    def modify_config(cls, config):
        config.relaxation_enabled = False

    @classmethod
    def update_defaults(cls, defaults):
        defaults.update({'lat_nx': 64, 'lat_ny': 66, 'lat_nz': 128, 'grid': 'D3Q19', 'subdomains': 2})


GEO = EqualSubdomainsGeometry3D

# subdomain id -> [(index tuple, value)]           (3d_propagation.py:128-146)
PLACED = {
    0: [(p2(vi(*p1(1, 0, 0)), 32, 32, 64), 0.11), (p2(vi(*p1(1, 1, 0)), 32, 32, 64), 0.12),
        (p2(vi(*p1(1, -1, 0)), 32, 32, 64), 0.13), (p2(vi(*p1(1, 0, 1)), 32, 32, 64), 0.21),
        (p2(vi(*p1(1, 0, -1)), 32, 32, 64), 0.23)],
    1: [(p2(vi(*p1(-1, 0, 0)), 1, 1, 1), 0.33), (p2(vi(*p1(-1, 1, 0)), 1, 1, 1), 0.34),
        (p2(vi(*p1(-1, 0, 1)), 1, 1, 1), 0.35),
        (p2(vi(*p1(-1, 0, 0)), 1, 32, 1), 0.44), (p2(vi(*p1(-1, 1, 0)), 1, 32, 1), 0.45),
        (p2(vi(*p1(-1, -1, 0)), 1, 32, 1), 0.46), (p2(vi(*p1(-1, 0, 1)), 1, 32, 1), 0.47)],
}
# after one AB step                                 (3d_propagation.py:150-169)
EXPECTED = {
    1: [(p2(vi(*p1(1, 0, 0)), 32, 32, 1), 0.11), (p2(vi(*p1(1, 1, 0)), 32, 33, 1), 0.12),
        (p2(vi(*p1(1, -1, 0)), 32, 31, 1), 0.13), (p2(vi(*p1(1, 0, 1)), 33, 32, 1), 0.21),
        (p2(vi(*p1(1, 0, -1)), 31, 32, 1), 0.23)],
    0: [(p2(vi(*p1(-1, 0, 0)), 1, 1, 64), 0.33), (p2(vi(*p1(-1, 1, 0)), 1, 2, 64), 0.34),
        (p2(vi(*p1(-1, 0, 1)), 2, 1, 64), 0.35),
        (p2(vi(*p1(-1, 0, 0)), 1, 32, 64), 0.44), (p2(vi(*p1(-1, 1, 0)), 1, 33, 64), 0.45),
        (p2(vi(*p1(-1, -1, 0)), 1, 31, 64), 0.46), (p2(vi(*p1(-1, 0, 1)), 2, 32, 64), 0.47)],
}

# AA analogue of AAPropagationTest.test_4subdomains (:707-721) for a z-split: after the first
# (in-place) step nothing has streamed, the value sits in the OPPOSITE slot and has been copied
# into the neighbour's ghost layer; after the second (odd) step it has streamed TWICE (the AA
# pattern moves populations two nodes every second step; cf. test_2subdomains :723-733, where
# (z, y) = (2, 37) -> (40, 39) after two steps with periodic z).
AA_PLACED = {0: [((vi(0, 1, 1), 64, 32, 10), 0.11)]}
AA_AFTER_1 = {1: [((vi(0, -1, -1), 0, 32, 10), 0.11)], 0: [((vi(0, -1, -1), 64, 32, 10), 0.11)]}
AA_AFTER_2 = {1: [((vi(0, 1, 1), 2, 34, 10), 0.11)]}
