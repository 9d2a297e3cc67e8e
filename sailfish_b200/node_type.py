"""Node (boundary-condition) types of the hot path, mirroring sailfish/node_type.py.

Same class names, constructor arguments and flags as the reference so that user
``Subdomain.boundary_conditions`` code is source compatible.  Only the node types the
B200 kernels implement exist here; the numeric ``id`` values are the ones the reference
assigns (alphabetical order over all its node classes, node_type.py:404-420), which
keeps the per-run id remapping of the encoder (geo_encoder.py:83-91) -- and therefore
the encoded node map -- identical.  Pinned by tests/test_host_geometry.py.
"""


class LBNodeType(object):
    id = None
    wet_node = False
    excluded = False
    propagation_only = False
    standard_macro = False
    needs_orientation = False
    link_tags = False
    location = 0.0
    allow_unused = False

    def __init__(self, **params):
        if 'orientation' in params:
            self.orientation = params.pop('orientation')
        self.params = params


class _NTFluid(LBNodeType):
    wet_node = True
    standard_macro = True
    id = 0


class _NTGhost(LBNodeType):
    excluded = True
    id = 20


class _NTPropagationOnly(LBNodeType):
    propagation_only = True
    id = 21


class _NTUnused(LBNodeType):
    excluded = True
    id = 22


class NTCopy(LBNodeType):
    """node_type.py:310-317: the populations that would arrive from outside are copied from the inward neighbour (a
    crude vanishing-gradient outflow); AB access pattern only, as in the reference."""
    wet_node = True
    standard_macro = True
    needs_orientation = True
    id = 1


class NTDoNothing(LBNodeType):
    """node_type.py:296-307."""
    wet_node = True
    needs_orientation = True
    standard_macro = True
    id = 2


class NTEquilibriumDensity(LBNodeType):
    """node_type.py:200-208."""
    needs_orientation = True
    wet_node = True
    id = 3

    def __init__(self, density, orientation=None):
        self.params = {'density': density}
        self.orientation = orientation


class NTEquilibriumVelocity(LBNodeType):
    """node_type.py:246-254."""
    needs_orientation = True
    wet_node = True
    id = 4

    def __init__(self, velocity, orientation=None):
        self.params = {'velocity': velocity}
        self.orientation = orientation


class NTFullBBWall(LBNodeType):
    """node_type.py:148-170."""
    standard_macro = True
    location = 0.5
    needs_orientation = True
    id = 6


class NTHalfBBWall(LBNodeType):
    """node_type.py:117-145."""
    wet_node = True
    standard_macro = True
    needs_orientation = True
    link_tags = True
    location = -0.5
    allow_unused = True
    id = 9


class NTRegularizedDensity(LBNodeType):
    """node_type.py:211-223."""
    needs_orientation = True
    wet_node = True
    id = 12

    def __init__(self, density, orientation=None):
        self.params = {'density': density}
        self.orientation = orientation


class NTRegularizedVelocity(LBNodeType):
    """node_type.py:269-281."""
    needs_orientation = True
    wet_node = True
    id = 13

    def __init__(self, velocity, orientation=None):
        self.params = {'velocity': velocity}
        self.orientation = orientation


class NTSlip(LBNodeType):
    """node_type.py:402-404: full-slip (specular reflection) node; not a wet node, and -- as in the reference -- its
    orientation is not detected: pass NTSlip(orientation=grid.vec_to_dir(inward normal))."""
    standard_macro = True
    id = 14


class NTWallTMS(LBNodeType):
    """node_type.py:171-192: wall for turbulent flows after Chikatamarla and Karlin (Tamm-Mott-Smith approximation):
    the populations coming out of the wall are the equilibrium at target moments (those of the bounced-back state),
    and f^eq(target) - f^eq(instantaneous) is added after the collision."""
    wet_node = True
    needs_orientation = True
    link_tags = True
    location = 0.5
    allow_unused = True
    standard_macro = True
    id = 15


class NTYuOutflow(LBNodeType):
    """node_type.py:336-350: open boundary of Yu, Mei and Shyy, f_i(x) = 2 f_i(x - n) - f_i(x - 2n) for the
    populations that would arrive from outside; AB access pattern only."""
    wet_node = True
    standard_macro = True
    needs_orientation = True
    id = 16


class NTZouHeDensity(LBNodeType):
    """node_type.py:229-239: imposed density, bounce-back of the non-equilibrium parts + tangential momentum fix-up."""
    needs_orientation = True
    wet_node = True
    id = 17

    def __init__(self, density, orientation=None):
        self.params = {'density': density}
        self.orientation = orientation


class NTZouHeVelocity(LBNodeType):
    """node_type.py:256-266."""
    needs_orientation = True
    wet_node = True
    id = 18

    def __init__(self, velocity, orientation=None):
        self.params = {'velocity': velocity}
        self.orientation = orientation


_NODE_TYPES = dict((c.id, c) for c in (
    _NTFluid, _NTGhost, _NTPropagationOnly, _NTUnused, NTCopy, NTYuOutflow, NTWallTMS, NTSlip, NTDoNothing, NTEquilibriumDensity,
    NTEquilibriumVelocity, NTFullBBWall, NTHalfBBWall, NTRegularizedDensity,
    NTRegularizedVelocity, NTZouHeDensity, NTZouHeVelocity))


def get_wet_node_type_ids(allow_unused=None):
    return [i for i, c in _NODE_TYPES.items()
            if c.wet_node and (allow_unused is None or c.allow_unused == allow_unused)]


def get_dry_node_type_ids():
    return [i for i, c in _NODE_TYPES.items() if not c.wet_node]


def get_orientation_node_type_ids():
    return [i for i, c in _NODE_TYPES.items() if c.needs_orientation]


def get_link_tag_node_type_ids():
    return [i for i, c in _NODE_TYPES.items() if c.link_tags]
