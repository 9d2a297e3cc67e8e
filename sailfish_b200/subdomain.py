"""Host-side geometry of a subdomain: node-type map, orientation / link tags, encoding.

Mirror of the parts of sailfish/subdomain.py (SubdomainSpec :32-347, Subdomain :350-870,
Subdomain2D/3D :872-1013) and sailfish/geo_encoder.py (GeoEncoderConst :52-383) that feed
the collide-and-propagate kernels.  User code written against the reference
(``boundary_conditions(self, hx, hy[, hz])`` calling ``self.set_node(where, NT...)``,
``initial_conditions(self, sim, hx, hy[, hz])``) runs unchanged; the encoded uint32 map,
``node_params`` and bit layout are bit-identical to the reference's
(tests/test_host_geometry.py compares against maps produced by the reference itself).
"""
import inspect

import numpy as np
from scipy.ndimage import convolve

from sailfish_b200 import node_type as nt


def bit_len(num):
    """geo_encoder.py:13-19."""
    length = 0
    while num:
        num >>= 1
        length += 1
    return max(length, 1)


def is_number(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


class SubdomainSpec(object):
    """Location/size/neighbours of a subdomain (sailfish/subdomain.py:32-347, reduced to
    what a split along one axis needs)."""
    X_LOW, X_HIGH, Y_LOW, Y_HIGH, Z_LOW, Z_HIGH = range(6)

    def __init__(self, location, size, envelope_size=1, id_=0):
        self.location = tuple(location)
        self.size = tuple(size)
        self.dim = len(self.size)
        self.envelope_size = envelope_size
        self.actual_size = tuple(s + 2 * envelope_size for s in self.size)
        self.id = id_
        self._periodicity = [False] * self.dim
        self._face_conn = {}        # face -> neighbour subdomain id

    ox = property(lambda self: self.location[0])
    oy = property(lambda self: self.location[1])
    oz = property(lambda self: self.location[2])
    nx = property(lambda self: self.size[0])
    ny = property(lambda self: self.size[1])
    nz = property(lambda self: self.size[2])

    @property
    def num_nodes(self):
        return int(np.prod(self.size))

    def enable_local_periodicity(self, axis):
        self._periodicity[axis] = True

    def has_face_conn(self, face):
        return face in self._face_conn

    def connect(self, face, other_id):
        self._face_conn[face] = other_id

    @property
    def _nonghost_slice(self):
        es = self.envelope_size
        return tuple(slice(es, es + n) for n in reversed(self.size))


class Subdomain(object):
    """sailfish/subdomain.py:350-870."""
    dim = None

    def __init__(self, grid_shape, spec, grid, config, mem_alignment=32):
        self.spec = spec
        self.grid_shape = tuple(grid_shape)
        self.grid = grid
        self.config = config
        es = spec.envelope_size
        lat_shape = tuple(reversed(spec.actual_size))
        arr_nx = int(np.ceil(lat_shape[-1] / float(mem_alignment))) * mem_alignment
        self.lat_shape = lat_shape
        self.arr_shape = lat_shape[:-1] + (arr_nx,)
        # base arrays cover ghosts and x padding (subdomain_runner.py:253-320, subdomain.py:396-404)
        self._type_map_base = np.zeros(self.arr_shape, dtype=np.uint32)
        self._param_map_base = np.zeros(self.arr_shape, dtype=np.intp)
        self._orientation_base = np.zeros(self.arr_shape, dtype=np.uint32)
        self._params = {}
        self._seen_types = set([0])
        self._needs_orientation = False
        self._type_map_encoded = False
        self._type_vis_map = None
        self.node_params = []
        self.type_id_remap = {0: 0}
        self.bits = (1, 1, 0)
        self._es = es

    # -- API used by user code ----------------------------------------------------------
    def boundary_conditions(self, *args):
        raise NotImplementedError('boundary_conditions() not defined in a child class.')

    def initial_conditions(self, sim, *args):
        raise NotImplementedError('initial_conditions() not defined in a child class')

    @staticmethod
    def _hashable_params(param_dict):
        out = []
        for k, v in param_dict.items():
            out.append((k, v.tobytes() if hasattr(v, 'tobytes') else v))
        return frozenset(out)

    def set_node(self, where, node_type):
        """subdomain.py:532-562."""
        assert not self._type_map_encoded
        # one index computation shared by the assignments below (flat indices: cheaper than a tuple of
        # per-axis index arrays on a 10^8-node map)
        base = self._type_map_base
        where_idx = np.flatnonzero(np.broadcast_to(np.asarray(where, dtype=bool), base.shape))
        # the reference's base arrays have no x padding (subdomain.py:887-905): a condition on y / z alone must not
        # select padding columns here either
        if base.shape[-1] != self.lat_shape[-1]:
            where_idx = where_idx[(where_idx % base.shape[-1]) < self.lat_shape[-1]]
        if inspect.isclass(node_type):
            assert issubclass(node_type, nt.LBNodeType)
            node_type = node_type()
        else:
            assert isinstance(node_type, nt.LBNodeType)
        for name, param in node_type.params.items():
            if is_number(param):
                continue
            if type(param) is tuple and all(is_number(e) for e in param):
                continue
            raise ValueError('Unsupported node parameter %r (%s): the B200 hot path takes numbers '
                             'and constant vectors' % (name, type(param).__name__))
        base.reshape(-1)[where_idx] = node_type.id
        key = hash((node_type.id, self._hashable_params(node_type.params)))
        pm = self._param_map_base.reshape(-1)
        assert not pm[where_idx].any(), 'Overriding previously set nodes is not allowed.'
        pm[where_idx] = key
        self._params[key] = node_type
        self._seen_types.add(node_type.id)
        if getattr(node_type, 'orientation', None) is not None:
            self._orientation_base.reshape(-1)[where_idx] = node_type.orientation
        elif node_type.needs_orientation:
            self._needs_orientation = True

    # -- geometry processing ---------------------------------------------------------------
    def _lat_view(self, arr):
        return arr[..., :self.lat_shape[-1]]

    @staticmethod
    def _member(type_map, ids):
        """np.isin(type_map, ids) for small non-negative type ids, as a table look-up (one pass)."""
        ids = [int(i) for i in ids]
        lut = np.zeros(max([int(type_map.max()) if type_map.size else 0] + ids) + 1, dtype=bool)
        lut[ids] = True
        return lut[type_map]

    def _fluid_map(self, wet, allow_unused=None):
        src = self._lat_view(self._type_map_base)
        if wet:
            return self._member(src, nt.get_wet_node_type_ids(allow_unused=allow_unused))
        return src == 0

    def _lattice_kernel(self, zero=1):
        k = np.zeros((3,) * self.dim, dtype=np.uint8)
        k[(1,) * self.dim] = zero
        for e in self.grid.basis:
            k[tuple(1 + int(c) for c in reversed(e))] = 1
        return k

    def _postprocess_nodes(self):
        """subdomain.py:845-869."""
        tm = self._lat_view(self._type_map_base)
        fluid_map = self._fluid_map(wet=False).astype(np.uint8)
        wet_unused = self._fluid_map(wet=True, allow_unused=True)
        wet_map = self._fluid_map(wet=True).astype(np.uint8)
        nb = self._lattice_kernel()
        where = convolve(fluid_map, nb, mode='constant', cval=1) == 0
        tm[where & wet_unused] = nt._NTUnused.id
        where = convolve(wet_map, nb, mode='constant', cval=0) == 0
        tm[where & np.logical_not(wet_map.astype(bool))] = nt._NTUnused.id
        used = (tm != nt._NTUnused.id).astype(np.uint8)
        where = convolve(used, nb, mode='constant', cval=0) > 0
        tm[where & (tm == nt._NTUnused.id)] = nt._NTPropagationOnly.id

    def _tag_directions(self):
        """subdomain.py:593-642."""
        tm_full = self._lat_view(self._type_map_base)
        ob_full = self._lat_view(self._orientation_base)
        ngs = list(self.spec._nonghost_slice)
        for i, periodic in enumerate(reversed(self.spec._periodicity)):
            if not periodic:
                ngs[i] = slice(None)
        ngs = tuple(ngs)
        uniq = set(np.unique(tm_full).tolist())
        wet = sorted(set(nt.get_wet_node_type_ids()) & uniq)
        orient_types = sorted(set(nt.get_link_tag_node_type_ids()) & uniq)
        if not orient_types:
            return False
        tm = tm_full[ngs]
        ob = ob_full[ngs]
        orient_map = self._member(tm, orient_types) & (ob == 0)
        l = self.dim - 1
        for i, vec in enumerate(self.grid.basis[1:]):
            shifted = tm
            for j, shift in enumerate(vec):
                if shift == 0:
                    continue
                shifted = np.roll(shifted, int(-shift), axis=l - j)
            idx = orient_map & self._member(shifted, wet)
            ob[idx] |= np.uint32(1 << i)
        return True

    def _detect_orientation(self, use_tags):
        """subdomain.py:644-673."""
        tm = self._lat_view(self._type_map_base)
        ob = self._lat_view(self._orientation_base)
        uniq = set(np.unique(tm).tolist())
        skip = set(nt.get_link_tag_node_type_ids()) if use_tags else set()
        orient_types = sorted((set(nt.get_orientation_node_type_ids()) - skip) & uniq)
        if not orient_types:
            return
        orient_map = np.isin(tm, orient_types)
        l = self.dim - 1
        for vec in self.grid.basis:
            if int(np.dot(vec, vec)) != 1:
                continue
            shifted = tm
            for j, shift in enumerate(vec):
                if shift == 0:
                    continue
                shifted = np.roll(shifted, int(-shift), axis=l - j)
            idx = orient_map & (shifted == 0) & (ob == 0)
            ob[idx] = self.grid.vec_to_dir(vec)

    def _define_ghosts(self, unset_only=False):
        """subdomain.py:907-940 / :978-1005 (split faces keep the exchanged nodes fluid
        until orientation detection is done)."""
        es = self._es
        tm = self._type_map_base
        spec = self.spec
        inner = (slice(es, -es),) * (self.dim - 1)

        def _set(view, face):
            if unset_only and spec.has_face_conn(face):
                tg = view == 0
                # nodes exchanged with the neighbour stay as they are
                keep = np.zeros(view.shape, bool)
                sl = [slice(None)] * view.ndim
                # the face view has thickness es along its normal axis; the exchanged
                # rectangle is the real cross-section
                axis = self.dim - 1 - face // 2
                for a in range(view.ndim):
                    if a != axis:
                        n = self.lat_shape[a]
                        sl[a] = slice(es, n - es)
                keep[tuple(sl)] = True
                tg &= ~keep
                view[tg] = nt._NTGhost.id
            else:
                view[...] = nt._NTGhost.id
        lx = self.lat_shape[-1]
        if self.dim == 2:
            _set(tm[0:es, :], spec.Y_LOW)
            _set(tm[:, 0:es], spec.X_LOW)
            _set(tm[es + spec.ny:, :], spec.Y_HIGH)
            _set(tm[:, es + spec.nx:], spec.X_HIGH)
        else:
            _set(tm[0:es, :, :], spec.Z_LOW)
            _set(tm[:, 0:es, :], spec.Y_LOW)
            _set(tm[:, :, 0:es], spec.X_LOW)
            _set(tm[es + spec.nz:, :, :], spec.Z_HIGH)
            _set(tm[:, es + spec.ny:, :], spec.Y_HIGH)
            _set(tm[:, :, es + spec.nx:], spec.X_HIGH)
        del lx, inner

    def reset(self):
        """subdomain.py:675-737."""
        self._type_map_encoded = False
        self.boundary_conditions(*self._get_mgrid_base())
        have_link_tags = False
        self._define_ghosts(unset_only=True)
        self._postprocess_nodes()
        use_tags = getattr(self.config, 'use_link_tags', True)
        if self._needs_orientation:
            if use_tags:
                have_link_tags = self._tag_directions()
            self._detect_orientation(use_tags)
        self._define_ghosts()
        self._type_vis_map = self._type_map_base[self.spec._nonghost_slice].copy()
        self._encode(have_link_tags)

    # -- encoding: geo_encoder.py:73-383 --------------------------------------------------------
    def _encode(self, have_link_tags):
        tm = self._type_map_base
        uniq = [int(x) for x in np.unique(tm)]
        self.node_types = set([nt._NTFluid] + [nt._NODE_TYPES[i] for i in uniq])
        remap = {0: 0}
        for i, t in enumerate(uniq):
            remap[t] = i + 1
        self.type_id_remap = remap
        bits_type = bit_len(len(uniq))
        encoded_param = np.zeros_like(tm)
        geo_params = []
        param_to_idx = {}
        items = 0
        for key, node_type in self._params.items():
            for param in node_type.params.values():
                if is_number(param):
                    if param in param_to_idx:
                        idx = param_to_idx[param]
                    else:
                        geo_params.append(param)
                        idx = items
                        param_to_idx[param] = idx
                        items += 1
                else:
                    if param in param_to_idx:
                        idx = param_to_idx[param]
                    else:
                        geo_params.extend(param)
                        idx = items
                        param_to_idx[param] = idx
                        items += len(param)
                encoded_param[self._param_map_base == key] = idx
        bits_param = bit_len(items)
        self.node_params = [float(x) for x in geo_params]
        self.bits = (bits_type, bits_param, 0)
        choice = np.zeros(max(remap.keys()) + 1, dtype=np.uint32)
        for k, v in remap.items():
            choice[k] = v
        misc = self._orientation_base.astype(np.uint32)
        misc = (misc << np.uint32(bits_param)) | encoded_param.astype(np.uint32)
        self._encoded = ((misc << np.uint32(bits_type)) | choice[tm]).astype(np.uint32)
        self._type_map_encoded = True
        self.have_link_tags = have_link_tags
        # geo_encoder.py:283-289: tag bits that are set in every tagged node
        self.unused_tag_bits = 0
        if have_link_tags:
            tags = self._orientation_base[self._orientation_base > 0]
            if tags.size:
                self.unused_tag_bits = int(np.bitwise_and.reduce(tags.astype(np.uint32)))

    def encoded_map(self):
        return self._encoded

    def visualization_map(self):
        return self._type_vis_map

    def fluid_map(self, wet=True):
        """Wet-node mask over the non-ghost region (subdomain.py:797-809)."""
        fm = self._type_vis_map
        if wet:
            return np.isin(fm, nt.get_wet_node_type_ids())
        return fm == 0

    @property
    def num_fluid_nodes(self):
        return int(np.sum(self.fluid_map()))

    def update_context(self, ctx):
        """The encoder's share of the kernel configuration (geo_encoder.py:341-363 via subdomain.py:781-793)."""
        ctx.update({
            'use_link_tags': getattr(self.config, 'use_link_tags', True),
            'node_types': self.node_types,
            'type_id_remap': self.type_id_remap,
            'nt_misc_shift': self.bits[0],
            'nt_param_shift': self.bits[1],
            'nt_scratch_shift': self.bits[2],
            'node_params': self.node_params,
            'scratch_space': False,
            'unused_tag_bits': getattr(self, 'unused_tag_bits', 0),
        })

    def type_ids_by_name(self):
        return dict((c.__name__, int(self.type_id_remap[c.id])) for c in self.node_types)

    # -- coordinates ---------------------------------------------------------------------------
    def _periodic_flags(self):
        c = self.config
        return [bool(getattr(c, 'periodic_' + a, False)) for a in 'xyz'[:self.dim]]

    def _get_mgrid(self):
        """Index arrays (natural order hx, hy[, hz]) of the non-ghost region
        (subdomain.py:881-885 / :947-951).  Read-only broadcast views: same values and shape as
        np.mgrid without materialising dim x N int64 arrays."""
        s = self.spec
        if self.dim == 2:
            ax = [np.arange(s.oy, s.oy + s.ny), np.arange(s.ox, s.ox + s.nx)]
        else:
            ax = [np.arange(s.oz, s.oz + s.nz), np.arange(s.oy, s.oy + s.ny), np.arange(s.ox, s.ox + s.nx)]
        return tuple(reversed(self._broadcast(ax)))

    @staticmethod
    def _broadcast(axes):
        shape = tuple(a.size for a in axes)
        out = []
        for k, a in enumerate(axes):
            sh = [1] * len(axes)
            sh[k] = a.size
            out.append(np.broadcast_to(a.reshape(sh), shape))
        return out

    def _get_mgrid_base(self):
        """Global coordinates of every node incl. ghosts and x padding (the reference's base
        fields are padded too: subdomain.py:887-905 / :953-976 index them with np.where on the
        ghost-inclusive grid, so padding is simply never selected; here the padding columns get
        coordinates far outside the domain for the same effect).  Periodic axes wrap."""
        s, es = self.spec, self._es
        per = self._periodic_flags()
        shape = self.arr_shape
        if self.dim == 2:
            origins, sizes = [s.oy, s.ox], [self.gy, self.gx]
        else:
            origins, sizes = [s.oz, s.oy, s.ox], [self.gz, self.gy, self.gx]
        axes = []
        for k in range(self.dim):
            c = np.arange(origins[k] - es, origins[k] - es + shape[k])
            if list(reversed(per))[k]:
                n = sizes[k]
                c = np.where(c < 0, c + n, c)
                c = np.where(c >= n, c - n, c)
            if k == self.dim - 1:
                c[self.lat_shape[-1]:] = -(10 ** 9)     # x padding
            axes.append(c)
        return tuple(reversed(self._broadcast(axes)))

    def init_fields(self, sim):
        self.initial_conditions(sim, *self._get_mgrid())


class Subdomain2D(Subdomain):
    dim = 2

    def __init__(self, grid_shape, spec, grid, config, **kw):
        self.gy, self.gx = grid_shape
        Subdomain.__init__(self, grid_shape, spec, grid, config, **kw)


class Subdomain3D(Subdomain):
    dim = 3

    def __init__(self, grid_shape, spec, grid, config, **kw):
        self.gz, self.gy, self.gx = grid_shape
        Subdomain.__init__(self, grid_shape, spec, grid, config, **kw)


__all__ = ['SubdomainSpec', 'Subdomain', 'Subdomain2D', 'Subdomain3D', 'lat']
