"""ctypes mirror of ``SlfParams`` (include/slf_b200.h) and its construction from the
code-generation context of the reference.

The reference bakes these values into the generated CUDA source as literals
(codegen.py:195-204 ``_build_context``); here they are data handed to
``slf_plan_create``.  ``from_context`` accepts the very dict the reference builds
(sim.update_context + runner.update_context), so the plugin backend needs nothing
else; the self-contained host mirror in this package builds the same dict.
"""
import ctypes

import numpy as np

LATTICE_IDS = {'D2Q9': 0, 'D3Q19': 1, 'D3Q27': 2}
MODEL_IDS = {'bgk': 0, 'mrt': 1}
ACCESS_IDS = {'AB': 0, 'AA': 1}
SIMTYPE_IDS = {'lbm': 0, 'shan-chen': 1}
SC_POTENTIAL_IDS = {'linear': 0, 'classic': 1}

# Order must match the SLF_NT_* enum of include/slf_b200.h.
NODE_CLASS_NAMES = (
    '_NTFluid', '_NTGhost', '_NTUnused', '_NTPropagationOnly', 'NTFullBBWall',
    'NTHalfBBWall', 'NTEquilibriumVelocity', 'NTEquilibriumDensity',
    'NTRegularizedVelocity', 'NTRegularizedDensity', 'NTDoNothing', 'NTZouHeVelocity', 'NTZouHeDensity',
    'NTCopy', 'NTYuOutflow', 'NTWallTMS', 'NTSlip')
NT_COUNT = len(NODE_CLASS_NAMES)
NT_ABSENT = 0xfffffffe

OPTION_SAVE_MACRO_FIELDS = 1
OPTION_BULK = 2
OPTION_DENSITIES_ONLY = 4


class SlfParams(ctypes.Structure):
    _fields_ = [
        ('struct_bytes', ctypes.c_uint32),
        ('lattice', ctypes.c_int32),
        ('model', ctypes.c_int32),
        ('access_pattern', ctypes.c_int32),
        ('lat_nx', ctypes.c_int32), ('lat_ny', ctypes.c_int32), ('lat_nz', ctypes.c_int32),
        ('arr_nx', ctypes.c_int32), ('arr_ny', ctypes.c_int32), ('arr_nz', ctypes.c_int32),
        ('periodic', ctypes.c_int32 * 3),
        ('bits_type', ctypes.c_int32),
        ('bits_param', ctypes.c_int32),
        ('bits_scratch', ctypes.c_int32),
        ('type_id', ctypes.c_uint32 * NT_COUNT),
        ('use_link_tags', ctypes.c_int32),
        ('relaxation_enabled', ctypes.c_int32),
        ('propagation_enabled', ctypes.c_int32),
        ('force_enabled', ctypes.c_int32),
        ('unused_tag_bits', ctypes.c_int32),
        ('visc', ctypes.c_double),
        ('accel', ctypes.c_double * 3),
        ('num_node_params', ctypes.c_int32),
        ('simtype', ctypes.c_int32),
        ('node_params', ctypes.POINTER(ctypes.c_double)),
        ('tau_phi', ctypes.c_double),
        ('sc_g', ctypes.c_double * 4),
        ('sc_potential', ctypes.c_int32),
        ('reserved1', ctypes.c_int32),
        ('accel1', ctypes.c_double * 3),
    ]

    # keeps the node_params array alive
    _np_keepalive = None

    @property
    def dist_size(self):
        return self.arr_nx * self.arr_ny * self.arr_nz

    @property
    def dim(self):
        return 2 if self.lattice == 0 else 3

    @property
    def Q(self):
        return (9, 19, 27)[self.lattice]


class UnsupportedConfiguration(ValueError):
    """Raised for reference features outside the hot-path scope."""


def make_params(lattice, model, access_pattern, lat_size, arr_size, periodic,
                bits, type_ids, node_params, visc, accel=None, use_link_tags=True,
                relaxation_enabled=True, propagation_enabled=True, shan_chen=None, unused_tag_bits=0):
    """shan_chen: None (single fluid) or dict(tau_phi=, G11=, G12=, G21=, G22=, potential=)."""
    """lat_size/arr_size are (nx, ny, nz) with nz = 1 in 2-D; type_ids maps node
    class name -> remapped id."""
    p = SlfParams()
    p.struct_bytes = ctypes.sizeof(SlfParams)
    if lattice not in LATTICE_IDS:
        raise UnsupportedConfiguration('lattice %s (supported: %s)' % (lattice, ', '.join(sorted(LATTICE_IDS))))
    p.lattice = LATTICE_IDS[lattice]
    if model not in MODEL_IDS:
        raise UnsupportedConfiguration('relaxation model %s (supported: %s)' % (model, ', '.join(sorted(MODEL_IDS))))
    p.model = MODEL_IDS[model]
    p.access_pattern = ACCESS_IDS[access_pattern]
    p.lat_nx, p.lat_ny, p.lat_nz = [int(x) for x in lat_size]
    p.arr_nx, p.arr_ny, p.arr_nz = [int(x) for x in arr_size]
    for i in range(3):
        p.periodic[i] = int(bool(periodic[i]))
    p.bits_type, p.bits_param, p.bits_scratch = [int(b) for b in bits]
    for i, name in enumerate(NODE_CLASS_NAMES):
        p.type_id[i] = int(type_ids.get(name, NT_ABSENT))
    unknown = set(type_ids) - set(NODE_CLASS_NAMES)
    if unknown:
        raise UnsupportedConfiguration(
            'node types outside the B200 hot path: %s' % ', '.join(sorted(unknown)))
    if access_pattern == 'AA' and ('NTCopy' in type_ids or 'NTYuOutflow' in type_ids):
        # the reference generates '#error ... not supported with the AA access_pattern' (boundary.mako:626-628)
        raise UnsupportedConfiguration('NTCopy / NTYuOutflow nodes need --access_pattern=AB')
    p.use_link_tags = int(bool(use_link_tags))
    p.unused_tag_bits = int(unused_tag_bits) & 0x7fffffff
    p.relaxation_enabled = int(bool(relaxation_enabled))
    p.propagation_enabled = int(bool(propagation_enabled))
    p.visc = float(visc)
    if accel is not None and any(float(a) != 0.0 for a in accel):
        p.force_enabled = 1
        for i, a in enumerate(accel):
            p.accel[i] = float(a)
    arr = np.ascontiguousarray(np.asarray(list(node_params) or [0.0], dtype=np.float64))
    p._np_keepalive = arr
    p.num_node_params = int(arr.size)
    p.node_params = arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if shan_chen is not None:
        p.simtype = SIMTYPE_IDS['shan-chen']
        p.tau_phi = float(shan_chen.get('tau_phi', 1.0))
        for i, k in enumerate(('G11', 'G12', 'G21', 'G22')):
            p.sc_g[i] = float(shan_chen.get(k, 0.0))
        for i, a in enumerate(shan_chen.get('accel1') or ()):
            p.accel1[i] = float(a)
        p.sc_potential = SC_POTENTIAL_IDS[shan_chen.get('potential', 'linear')]
    return p


def from_context(ctx):
    """Builds SlfParams from the reference's Mako context dict
    (lb_base.py:120-137, lb_single.py:52-70, subdomain_runner.py:182-248,
    geo_encoder.py:341-363)."""
    grid = ctx['grid']
    name = grid.__name__ if hasattr(grid, '__name__') else grid.name
    cfg = ctx.get('config')
    if ctx.get('node_addressing', 'direct') != 'direct':
        raise UnsupportedConfiguration('--node_addressing=indirect is out of scope')
    simtype = ctx.get('simtype', 'lbm')
    if simtype not in SIMTYPE_IDS:
        raise UnsupportedConfiguration('simtype %r' % simtype)
    for flag in ('incompressible', 'regularized'):
        if ctx.get(flag):
            raise UnsupportedConfiguration('--%s is out of scope' % flag)
    if cfg is not None and getattr(cfg, 'minimize_roundoff', False):
        raise UnsupportedConfiguration('--minimize_roundoff is out of scope')
    if ctx.get('subgrid', 'none') != 'none':
        raise UnsupportedConfiguration('subgrid models are out of scope')
    if ctx.get('propagate_on_read'):
        raise UnsupportedConfiguration('--propagate_on_read is out of scope')
    if ctx.get('symbol_idx_map') or ctx.get('timeseries_data'):
        raise UnsupportedConfiguration('time/space dependent node parameters are out of scope')
    if ctx.get('scratch_space'):
        raise UnsupportedConfiguration('node scratch space (Grad nodes) is out of scope')
    remap = ctx['type_id_remap']
    type_ids = {}
    for nt_class in ctx['node_types']:
        type_ids[nt_class.__name__] = int(remap[nt_class.id])
    accel = accel1 = None
    forces = ctx.get('forces')
    if forces is not None and not isinstance(forces, type(NotImplemented)):
        numeric = getattr(forces, 'numeric', None) or {}
        symbolic = getattr(forces, 'symbolic', None) or {}
        if symbolic:
            raise UnsupportedConfiguration('symbolic body forces are out of scope')
        if numeric:
            grids_ok = {0, 1} if simtype == 'shan-chen' else {0}
            if not set(numeric.keys()) <= grids_ok or any(set(v.keys()) != {True} for v in numeric.values()):
                raise UnsupportedConfiguration('only acceleration-type body forces (on grid 0; on both components '
                                               'of the binary Shan-Chen model)')
            if ctx.get('force_implementation', 'guo') != 'guo':
                raise UnsupportedConfiguration('only --force_implementation=guo')
            if 0 in numeric:
                accel = [float(a) for a in numeric[0][True]] + [0.0] * (3 - len(numeric[0][True]))
            if 1 in numeric:
                accel1 = [float(a) for a in numeric[1][True]] + [0.0] * (3 - len(numeric[1][True]))
    shan_chen = None
    if simtype == 'shan-chen':
        if len(ctx.get('grids', [])) != 2:
            raise UnsupportedConfiguration('single-component Shan-Chen is out of scope (binary only)')
        consts = ctx.get('constants', {})
        shan_chen = dict(tau_phi=ctx['tau_phi'], potential=ctx.get('sc_potential', 'linear'),
                         G11=consts.get('G11', 0.0), G12=consts.get('G12', 0.0),
                         G21=consts.get('G21', 0.0), G22=consts.get('G22', 0.0), accel1=accel1)
    elif ctx.get('force_couplings'):
        raise UnsupportedConfiguration('force couplings without the Shan-Chen model')
    dim = 2 if name == 'D2Q9' else 3
    lat = (ctx['lat_nx'], ctx['lat_ny'], ctx['lat_nz'] if dim == 3 else 1)
    arr = (ctx['arr_nx'], ctx['arr_ny'], ctx['arr_nz'] if dim == 3 else 1)
    return make_params(
        name, ctx['model'], ctx['access_pattern'], lat, arr, ctx['block_periodicity'],
        (ctx['nt_misc_shift'], ctx['nt_param_shift'], ctx['nt_scratch_shift']),
        type_ids, [float(x) for x in ctx['node_params']], ctx['visc'], accel,
        use_link_tags=ctx.get('use_link_tags', True),
        relaxation_enabled=ctx.get('relaxation_enabled', True),
        propagation_enabled=ctx.get('propagation_enabled', True), shan_chen=shan_chen,
        unused_tag_bits=ctx.get('unused_tag_bits', 0))
