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
    """geo.py:69-135: --subdomains N --conn_axis {x,y,z}.  The default axis is the outermost (slowest-varying)
    one, where a face message is a run of contiguous planes (z in 3-D, y in 2-D) and the step overlaps the exchange
    with the interior; other axes use the dimension-ordered face exchange."""
    @classmethod
    def add_options(cls, group):
        super(_EqualMixin, cls).add_options(group)
        group.add_argument('--subdomains', type=int, default=1, help='number of subdomains to use')
        group.add_argument('--conn_axis', type=str, default=None, choices=['x', 'y', 'z'],
                           help='axis along which the subdomains will be connected')

    def subdomains(self, n=None):
        n = self.config.subdomains if n is None else n
        axis = {'x': 0, 'y': 1, 'z': 2}.get(self.config.conn_axis, self.dim - 1)
        if axis >= self.dim:
            raise ValueError('--conn_axis=%s in a %d-D simulation' % (self.config.conn_axis, self.dim))
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


class _BlockMixin(object):
    """--blocks=BX,BY[,BZ]: a regular grid of blocks (the layouts of regtest/subdomains/{2d,3d}_ldc.py:
    2 x 2 and 2 x 2 x 2 blocks; the last block of every axis takes the remainder)."""
    @classmethod
    def add_options(cls, group):
        super(_BlockMixin, cls).add_options(group)
        group.add_argument('--blocks', type=str, default='', help='blocks per axis, e.g. 2,2,2')
        group.add_argument('--subdomains', type=int, default=1, help='(ignored when --blocks is given)')

    def subdomains(self, n=None):
        blocks = [int(b) for b in self.config.blocks.split(',')] if self.config.blocks else [1] * self.dim
        if len(blocks) != self.dim or min(blocks) < 1:
            raise ValueError('--blocks needs %d positive integers' % self.dim)
        size = self.gsize
        q = [size[a] // blocks[a] for a in range(self.dim)]
        specs = []
        import itertools
        for idx in itertools.product(*[range(b) for b in blocks]):
            loc = [idx[a] * q[a] for a in range(self.dim)]
            sz = [q[a] if idx[a] < blocks[a] - 1 else size[a] - q[a] * (blocks[a] - 1) for a in range(self.dim)]
            specs.append(SubdomainSpec(tuple(loc), tuple(sz), id_=len(specs)))
        return specs


class BlockSubdomainsGeometry2D(_BlockMixin, LBGeometry2D):
    pass


class BlockSubdomainsGeometry3D(_BlockMixin, LBGeometry3D):
    pass


class EqualSubdomainsGeometry2D(_EqualMixin, LBGeometry2D):
    pass


class EqualSubdomainsGeometry3D(_EqualMixin, LBGeometry3D):
    pass
