"""Global geometry -> subdomain list: mirror of sailfish/geo.py:13-135."""
from sailfish_b200.subdomain import SubdomainSpec


class LBGeometry(object):
    dim = None

    def __init__(self, config):
        self.config = config
        self.gx = config.lat_nx
        self.gy = config.lat_ny
        if self.dim == 3:
            self.gz = config.lat_nz

    @classmethod
    def add_options(cls, group):
        pass

    @property
    def gsize(self):
        return [self.gx, self.gy] + ([self.gz] if self.dim == 3 else [])

    def subdomains(self, n=None):
        return [SubdomainSpec((0,) * self.dim, tuple(self.gsize))]


class LBGeometry2D(LBGeometry):
    dim = 2

    @classmethod
    def add_options(cls, group):
        group.add_argument('--lat_nx', help='lattice width', type=int, default=0)
        group.add_argument('--lat_ny', help='lattice height', type=int, default=0)
        group.add_argument('--periodic_x', action='store_true', default=False)
        group.add_argument('--periodic_y', action='store_true', default=False)


class LBGeometry3D(LBGeometry):
    dim = 3

    @classmethod
    def add_options(cls, group):
        LBGeometry2D.add_options(group)
        group.add_argument('--lat_nz', help='lattice depth', type=int, default=0)
        group.add_argument('--periodic_z', action='store_true', default=False)


class _EqualMixin(object):
    """geo.py:69-135: --subdomains N --conn_axis {x,y,z}.  The B200 data plane exchanges
    ghost layers along the outermost (slowest-varying) axis only, where a face message is a
    run of contiguous planes: z in 3-D, y in 2-D."""
    @classmethod
    def add_options(cls, group):
        super(_EqualMixin, cls).add_options(group)
        group.add_argument('--subdomains', type=int, default=1, help='number of subdomains to use')
        group.add_argument('--conn_axis', type=str, default=None, choices=['x', 'y', 'z'],
                           help='axis along which the subdomains will be connected')

    def subdomains(self, n=None):
        n = self.config.subdomains if n is None else n
        axis = self.dim - 1
        want = {'x': 0, 'y': 1, 'z': 2}.get(self.config.conn_axis, axis)
        if n > 1 and want != axis:
            raise ValueError('B200 backend: subdomains are connected along the outermost axis (%s)'
                             % 'xyz'[axis])
        size = self.gsize
        total = size[axis]
        q, r = divmod(total, n)
        specs = []
        for i in range(n):
            ext = q if i < n - 1 else q + r     # geo.py:129-135: the last block takes the remainder
            loc = [0] * self.dim
            loc[axis] = i * q
            sz = list(size)
            sz[axis] = ext
            specs.append(SubdomainSpec(tuple(loc), tuple(sz), id_=i))
        return specs


class EqualSubdomainsGeometry2D(_EqualMixin, LBGeometry2D):
    pass


class EqualSubdomainsGeometry3D(_EqualMixin, LBGeometry3D):
    pass
